"""GPU probe of the tensor-core fringe-sum kernels against a complex128 torch evaluation of the
same sum (development tool; the parity tests proper are tests/test_gpu_parity.py)."""
import json
import math
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from bayeslim_b200 import ops, _lib  # noqa: E402


def make_case(na, ns, nfreq, seed=0, extent=300.0, allpairs=True):
    rng = np.random.default_rng(seed)
    antv = rng.uniform(-extent, extent, size=(na, 3))
    antv[:, 2] *= 0.01
    zen = np.degrees(np.arccos(rng.uniform(0.05, 1.0, ns)))
    az = rng.uniform(0, 360, ns)
    freqs = np.linspace(100e6, 200e6, max(nfreq, 2))[:nfreq]
    if allpairs:
        ii, jj = np.triu_indices(na, 1)
    else:
        ii, jj = np.triu_indices(na, 0)
    flip = rng.uniform(size=len(ii)) < 0.3
    i = np.where(flip, jj, ii)
    j = np.where(flip, ii, jj)
    return antv, zen, az, freqs, i, j


def reference(antv, geom, A_rm, freqs, i, j, dev):
    """V[b, f] = sum_s A[f, s] exp(2 pi i (r_j - r_i).shat nu / c) in complex128."""
    shat = geom.shat[:geom.ns[0], :3]
    r = torch.as_tensor(antv, device=dev)
    u = r @ shat.T                                        # (na, ns)
    out = []
    for f, nu in enumerate(freqs):
        E = torch.exp(2j * math.pi * u * (nu / 2.99792458e8))
        M = (E.conj() * A_rm[f].to(torch.float64)[None]) @ E.T   # (na, na): sum conj(E_i) A E_j
        out.append(M[torch.as_tensor(i, device=dev), torch.as_tensor(j, device=dev)])
    return torch.stack(out, dim=1)                        # (nbl, nf)


def run(na, ns, nfreq, dyn=1.0, time_it=False, seed=0, extent=300.0, unit_max=8192):
    dev = torch.device("cuda")
    ops.UNIT_MAX_SRC = unit_max
    antv, zen, az, freqs, i, j = make_case(na, ns, nfreq, seed, extent=extent)
    geom = ops.Geometry([torch.as_tensor(zen)], [torch.as_tensor(az)], dev)
    g = torch.Generator(device="cpu").manual_seed(seed)
    A_rm = torch.randn(nfreq, ns, generator=g).abs() * torch.exp(dyn * torch.randn(nfreq, ns, generator=g))
    A_rm = (A_rm * 3.7e3).to(dev, torch.float32)
    A = ops.pack_planes(geom, [A_rm[None].contiguous()])
    til = ops.AntTiling(i, j, na, dev)
    tc = ops.TcTiling(i, j, na, dev)
    f64 = torch.as_tensor(freqs, device=dev, dtype=torch.float64)
    antvecs = torch.as_tensor(antv, device=dev)
    V = ops.fringe_sum_ant(A, antvecs, til, geom, f64, nfreq, tc=tc)[0, :, 0]
    torch.cuda.synchronize()
    ref = reference(antv, geom, A_rm, freqs, i, j, dev)
    scale = ref.abs().max().item()
    err = (V.to(torch.complex128) - ref).abs()
    out = dict(na=na, ns=ns, nfreq=nfreq, nitems=tc.nitems, fill=tc.fill, extent=extent,
               unit_max=unit_max,
               relmax=err.max().item() / scale, relrms=(err.pow(2).mean().sqrt().item() / scale))
    if out["relmax"] > 1e-4:
        # diagnose: error per antenna pair block
        e2 = torch.zeros(tc.ldp, tc.ldp, device=dev, dtype=torch.float64)
        x = np.minimum(i, j)
        y = np.maximum(i, j)
        e2[torch.as_tensor(x, device=dev), torch.as_tensor(y, device=dev)] = err.max(dim=1).values / scale
        blk = e2.reshape(tc.ldp // 8, 8, tc.ldp // 8, 8).amax(dim=(1, 3))
        out["bad_blocks8"] = int((blk > 1e-4).sum().item())
        out["blocks8"] = int(blk.numel())
        rows = (e2.amax(dim=1) > 1e-4).nonzero().flatten().tolist()
        cols = (e2.amax(dim=0) > 1e-4).nonzero().flatten().tolist()
        out["bad_rows"] = rows[:40]
        out["bad_cols"] = cols[:40]
        out["sample"] = [[complex(a).real, complex(a).imag, complex(b).real, complex(b).imag]
                         for a, b in zip(V[:4, 0].tolist(), ref[:4, 0].tolist())]
    if time_it:
        if til.usable and not (len(sys.argv) > 1 and sys.argv[1] == "time"):
            V2 = ops.fringe_sum_ant(A, antvecs, til, geom, f64, nfreq)[0, :, 0]
            out["relmax_fp32"] = ((V2.to(torch.complex128) - ref).abs().max().item() / scale)
        for name, kw in (("tc", dict(tc=tc)), ("fp32", dict())):
            if name == "fp32" and (not til.usable or (len(sys.argv) > 1 and sys.argv[1] == "time")):
                continue
            for _ in range(2):
                ops.fringe_sum_ant(A, antvecs, til, geom, f64, nfreq, **kw)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            n = 3
            for _ in range(n):
                ops.fringe_sum_ant(A, antvecs, til, geom, f64, nfreq, **kw)
            torch.cuda.synchronize()
            out["ms_" + name] = (time.perf_counter() - t0) / n * 1e3
        out["evals"] = float(len(i)) * ns * nfreq
    return out


def run_bwd(na, ns, nfreq, need_r=True, seed=0, time_it=False):
    """Tensor-core backward against complex128 autograd of the same sum."""
    dev = torch.device("cuda")
    ops.UNIT_MAX_SRC = 8192
    antv, zen, az, freqs, i, j = make_case(na, ns, nfreq, seed)
    geom = ops.Geometry([torch.as_tensor(zen)], [torch.as_tensor(az)], dev)
    g = torch.Generator(device="cpu").manual_seed(seed)
    A_rm = (torch.randn(nfreq, ns, generator=g).abs() * 3.7e3).to(dev, torch.float32)
    G = torch.complex(torch.randn(len(i), nfreq, generator=g, dtype=torch.float64),
                      torch.randn(len(i), nfreq, generator=g, dtype=torch.float64)).to(dev) * 2.5e-3
    til = ops.AntTiling(i, j, na, dev)
    tc = ops.TcTiling(i, j, na, dev)
    f64 = torch.as_tensor(freqs, device=dev, dtype=torch.float64)
    out = dict(kind="bwd", na=na, ns=ns, nfreq=nfreq, need_r=need_r)
    res = {}
    for name, kw in (("tc", dict(tc=tc)), ("fp32", dict())):
        if name == "fp32" and not til.usable:
            continue
        Ain = A_rm.clone().requires_grad_(True)
        antvecs = torch.as_tensor(antv, device=dev).requires_grad_(need_r)
        A = ops.pack_planes(geom, [Ain[None]])
        V = ops.fringe_sum_ant(A, antvecs, til, geom, f64, nfreq, **kw)[0, :, 0]
        loss = torch.sum(G.real.float() * V.real + G.imag.float() * V.imag)
        loss.backward()
        res[name] = (Ain.grad.double(), antvecs.grad.double() if need_r else None)
        if time_it:
            Gc = torch.zeros(1, len(i), 1, nfreq, dtype=torch.complex64, device=dev)
            Gc[0, :, 0] = G.to(torch.complex64)
            fn = ops._tc_backward if name == "tc" else ops._ant_backward
            t_or_til = tc if name == "tc" else til
            antv4 = til.antv4(antvecs)
            for _ in range(2):
                fn(Gc, A.detach(), antv4, geom, f64, nfreq, 0, t_or_til, A.shape, True, need_r)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                fn(Gc, A.detach(), antv4, geom, f64, nfreq, 0, t_or_til, A.shape, True, need_r)
            torch.cuda.synchronize()
            out["ms_" + name] = (time.perf_counter() - t0) / 3 * 1e3
    if ns * na * nfreq <= 4e7:
        # complex128 autograd reference
        Aref = A_rm.double().clone().requires_grad_(True)
        r = torch.as_tensor(antv, device=dev).requires_grad_(True)
        shat = geom.shat[:ns, :3]
        u = r @ shat.T
        ii, jj = torch.as_tensor(i, device=dev), torch.as_tensor(j, device=dev)
        loss = 0
        for f, nu in enumerate(freqs):
            E = torch.exp(2j * math.pi * u * (nu / 2.99792458e8))
            M = (E.conj() * Aref[f][None]) @ E.T
            Vf = M[ii, jj]
            loss = loss + torch.sum(G[:, f].real * Vf.real + G[:, f].imag * Vf.imag)
        loss.backward()
        ref = (Aref.grad, r.grad)
    else:
        ref = res.get("fp32")
        out["ref"] = "fp32 kernels"
    for name, (dA, dr) in res.items():
        out["dA_relmax_" + name] = ((dA - ref[0]).abs().max() / ref[0].abs().max()).item()
        if need_r:
            out["dr_relmax_" + name] = ((dr - ref[1]).abs().max() / ref[1].abs().max()).item()
    return out


if __name__ == "__main__":
    cases = [(40, 128, 2, 1.0, False), (130, 640, 3, 1.0, False), (350, 4096, 2, 2.0, False),
             (350, 98304, 64, 1.0, True)]
    if len(sys.argv) > 1 and sys.argv[1] == "quick":
        cases = cases[:3]
    if len(sys.argv) > 1 and sys.argv[1] == "trunc":
        # accumulation-length study: coherent sums (tiny array) and random ones
        cases = [(350, 98304, 4, 1.0, False, 0, ext, um) for ext in (300.0, 2.0)
                 for um in (8192, 2048, 512, 128)]
    if len(sys.argv) > 1 and sys.argv[1] == "time":
        cases = [(350, 98304, 64, 1.0, True)]
    if len(sys.argv) > 1 and sys.argv[1] == "prof":
        cases = [(350, 98304, 16, 1.0, False)]
    if len(sys.argv) > 1 and sys.argv[1] == "all":
        cases = cases + [(350, 98304, 4, 1.0, False, 0, ext, 8192) for ext in (300.0, 2.0)]
    res = []
    if len(sys.argv) > 1 and sys.argv[1] == "profbwd":
        print(json.dumps(run_bwd(350, 98304, 16, True)), flush=True)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "bwd":
        for c in [(40, 128, 2, True), (130, 640, 3, True), (130, 640, 3, False),
                  (350, 4096, 2, True), (350, 4096, 2, False),
                  (350, 98304, 64, True, 0, True), (350, 98304, 64, False, 0, True)]:
            try:
                r = run_bwd(*c)
            except Exception as e:  # noqa: BLE001
                import traceback
                r = dict(case=list(c), error=repr(e), tb=traceback.format_exc()[-1500:])
            print(json.dumps(r), flush=True)
            if "error" in r:
                break
        sys.exit(0)
    for c in cases:
        try:
            r = run(*c)
        except Exception as e:  # noqa: BLE001
            r = dict(case=list(c), error=repr(e))
        print(json.dumps(r), flush=True)
        res.append(r)
        if "error" in r:
            break
    with open("gpurun_out/tc_probe.json", "w") as f:
        json.dump(res, f, indent=1)
