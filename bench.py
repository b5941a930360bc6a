#!/usr/bin/env python
"""bench.py -- layers/sec of the RCWA hot path on BASELINE.json configs[1]
(order 15x15, 1 patterned layer, 512-wavelength sweep, complex64 API) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--points P] [--order O]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the UNMODIFIED reference (baseline/_ref) on the host cores
    python bench.py --config 3|5 ...          # BASELINE configs[2] (order 21, 8 layers, complex128) / configs[4] (forward + backward)

One "step" = one chunk of P wavelengths of the 512-point sweep (Example1 cell, a-Si:H pillar with a
synthetic linear dispersion, SiO2 half space) through
    rcwa() -> add_input_layer -> set_incident_angle -> add_layer -> solve_global_smatrix -> S_parameters
i.e. Fourier factorisation -> eigendecomposition -> layer S-matrix -> Redheffer product with the
substrate -> t_xx(0,0).  Every step takes the next chunk of wavelengths (new inputs, working set
>> L2).  `value` times steps whose inputs are already resident in HBM; `e2e` times the same steps
fed from pinned HOST buffers with the S-parameters read back.  Multi-GPU: ranks take disjoint
contiguous slices of the sweep (weak scaling: P points per rank per step), the only collective is
the final all_gather of the S-parameters.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# a step allocates and frees ~10 buffers of 7.6 GB each: growable segments keep the caching allocator from fragmenting
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

import numpy as np  # noqa: E402
import torch  # noqa: E402

# oracle/ is imported only inside the CPU legs below (cpu_baseline / --impl reference); the B200 arm builds its
# synthetic inputs with the product package's own rasteriser.

N_SWEEP = 512
LAM0, LAM1 = 400.0, 700.0
# a-Si:H permittivity at 532 / 650 nm: cubic interpolation of the reference's example/Materials_data/aSiH.txt
SI_EPS = {532.0: complex(12.011610263133004, 0.5259120147560001), 650.0: complex(10.362267239174999, 0.15362360819199997)}


def eps_si(lam):
    """Synthetic linear dispersion through the two a-Si:H points above."""
    a, b = SI_EPS[532.0], SI_EPS[650.0]
    return a + (lam - 532.0) / (650.0 - 532.0) * (b - a)


def sweep_inputs(order):
    """Example1 cell (Example1.ipynb:40-57): 300 x 300 nm period, 180 x 100 nm pillar, 300 x 300 samples, glass substrate."""
    import torcwa_b200
    case = {"order": [order, order], "L": [300.0, 300.0], "eps_in": 1.46 ** 2}
    geo = torcwa_b200.geometry(Lx=300.0, Ly=300.0, nx=300, ny=300, edge_sharpness=1000.0, dtype=torch.float32, device=torch.device("cpu"))
    mask = geo.rectangle(Wx=180.0, Wy=100.0, Cx=150.0, Cy=150.0)
    lams = torch.linspace(LAM0, LAM1, N_SWEEP, dtype=torch.float32)
    return case, mask, lams


def make_grids(mask, lams):
    e = torch.tensor([eps_si(float(l)) for l in lams], dtype=torch.complex64)
    return mask[None] * e[:, None, None] + (1.0 - mask)[None]


# ------------------------------------------------------------------------------------------- the unmodified reference
def reference_module():
    """kch3782/torcwa as installed (unmodified) into baseline/_ref by `pip install --no-deps --target baseline/_ref
    /root/reference`; None if it is not there."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "torcwa")):
        return None
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        import torcwa
        return torcwa
    except Exception:
        return None


def reference_point(torcwa, order, cdtype, lam, device, mask=None):
    """One design point of the workload's unit (Example1 cell, one patterned layer + SiO2 half space) through the
    reference's own public API and stock code path; returns (seconds, t_xx)."""
    rd = torch.float32 if cdtype == torch.complex64 else torch.float64
    if mask is None:
        import torcwa_b200
        geo = torcwa_b200.geometry(Lx=300.0, Ly=300.0, nx=300, ny=300, edge_sharpness=1000.0, dtype=rd, device=torch.device("cpu"))
        mask = geo.rectangle(Wx=180.0, Wy=100.0, Cx=150.0, Cy=150.0).to(device)
    e = eps_si(float(lam))
    if device.type == "cuda":
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    sim = torcwa.rcwa(freq=1 / torch.tensor(float(lam), dtype=rd, device=device), order=[order, order], L=[300.0, 300.0], dtype=cdtype, device=device)
    sim.add_input_layer(eps=1.46 ** 2)
    sim.set_incident_angle(inc_ang=0.0, azi_ang=0.0)
    sim.add_layer(thickness=300.0, eps=mask * e + (1.0 - mask))
    sim.solve_global_smatrix()
    t = sim.S_parameters(orders=[0, 0], direction="forward", port="transmission", polarization="xx", ref_order=[0, 0])
    if device.type == "cuda":
        torch.cuda.synchronize()
    return time.perf_counter() - t0, complex(t.reshape(-1)[0]), mask


def cuda_baseline(order, points=8, warm=2):
    """The reference's PyTorch-CUDA path (unmodified baseline/_ref, device='cuda', allow_tf32=False as its README asks,
    sequential loop over wavelengths as its examples do) on the same GPU, in the same run: the denominator of the
    north star's '>= 10x the reference PyTorch-CUDA path' (BASELINE.md 3.2)."""
    torcwa = reference_module()
    if torcwa is None:
        return {"unavailable": "baseline/_ref is missing"}
    torch.backends.cuda.matmul.allow_tf32 = False
    dev = torch.device("cuda", torch.cuda.current_device())
    out = {"impl": "unmodified reference (baseline/_ref), device='cuda', allow_tf32=False", "warmups": warm, "points": points}
    lams = np.linspace(LAM0, LAM1, points + warm)
    for name, cd in (("c64", torch.complex64), ("c128", torch.complex128)):
        ts, mask = [], None
        try:
            for i, lam in enumerate(lams):
                dt, _, mask = reference_point(torcwa, order, cd, lam, dev, mask)
                if i >= warm:
                    ts.append(dt)
            out[name] = {"layers_per_s": len(ts) / sum(ts), "s_per_layer": sum(ts) / len(ts)}
        except Exception as ex:            # report, do not hide
            out[name] = {"error": str(ex)[:200]}
        torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------- B200 arm
def run_step(grids_dev, freq_dev, case, device, symmetry=None):
    """symmetry: None = the package default (auto-detected symmetry reduction on), False = the general path."""
    import torcwa_b200
    sim = torcwa_b200.rcwa(freq=freq_dev, order=case["order"], L=case["L"], dtype=torch.complex64, device=device, symmetry_reduction=symmetry)
    sim.add_input_layer(eps=case["eps_in"])
    sim.set_incident_angle(inc_ang=0.0, azi_ang=0.0)
    sim.add_layer(thickness=300.0, eps=grids_dev)
    sim.solve_global_smatrix()
    return sim.S_parameters(orders=[0, 0], direction="forward", port="transmission", polarization="xx", ref_order=[0, 0])


def count_my_launches(fn):
    """Kernels of librcwa_b200.so launched by one step (CUPTI through torch.profiler)."""
    mine = ("zgemm_grouped", "fill_strided", "dft_rows", "dft_cols", "toeplitz", "pq_assemble", "kz_branch", "layer_form", "layer_finish",
            "blockdiag_dense", "identity_kernel", "axpby", "lu_panel", "lu_perm", "lu_colswap", "tri_inv", "gather_cols", "eig_backward", "conj_transpose", "hess_step",
            "hess_fused", "hess_advance", "hb_col", "hb_matvec", "hb_zero", "bd_left_mul", "bd_right_mul", "bd_add", "qr_pass", "qr_init", "qr_count", "qr_finish", "qr_stats", "diag_extract", "tnorm", "trevc_block",
            "colnorm", "colscale", "sym_project", "tc_gemm_kernel", "tc_split_rows", "tc_split_cols", "tc_colmax", "tc_fill_int", "tc_fix_exponent")
    try:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            fn()
            torch.cuda.synchronize()
        per = {}
        for e in prof.events():
            nm = e.name
            for m in mine:
                if m in nm:
                    d = per.setdefault(m, [0, 0.0])
                    d[0] += 1
                    d[1] += float(getattr(e, "device_time", 0.0) or getattr(e, "cuda_time", 0.0) or 0.0)
                    break
        return sum(v[0] for v in per.values()), per
    except Exception as ex:        # CUPTI missing: report honestly that it was not counted
        return None, {"error": str(ex)}


def bench_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (args.gpus, world))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        # NCCL prints its version banner / debug lines to stdout by default: keep stdout for the ONE JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    case, mask, lams = sweep_inputs(args.order)
    P = args.points
    # Weak scaling: every rank solves P wavelengths per step.  Chunk c of the sweep = wavelengths
    # [c*P, (c+1)*P) mod 512; in step s rank r takes chunk (s*world + r), so ranks and steps never share
    # inputs until the 512-point sweep wraps around.
    grids_host = make_grids(mask, lams).pin_memory()                     # [512,300,300] c64, pinned
    freq_host = (1.0 / lams).pin_memory()
    grids_res = grids_host.to(device)                                    # resident copy for the device-timed leg
    freq_res = freq_host.to(device)

    def chunk(s):
        c = s * world + rank
        return (torch.arange(P) + c * P) % N_SWEEP

    def step_resident(s, symmetry=None):
        sl = chunk(s).to(device)
        return run_step(grids_res[sl], freq_res[sl], case, device, symmetry)

    def step_e2e(s):
        sl = chunk(s)
        g = grids_host[sl].pin_memory().to(device, non_blocking=True)     # host gather of this step's wavelengths, then H2D
        f = freq_host[sl].pin_memory().to(device, non_blocking=True)
        out = run_step(g, f, case, device)
        if world > 1:
            full = torch.empty((world,) + tuple(torch.view_as_real(out).shape), dtype=torch.float32, device=device)
            dist.all_gather_into_tensor(full, torch.view_as_real(out).contiguous())
            out = full
        return out.cpu()                                                  # device->host read of the step's result

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K, W, s0=0):
        for s in range(W):
            fn(s0 + s)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(K):
            fn(s0 + W + s)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms

    K, W = args.steps, args.warmup
    clocks = ClockSampler(local)
    clocks.start()
    ms_res = timed(step_resident, K, W)
    clk = clocks.stop()
    ms_e2e = timed(step_e2e, K, 1, s0=K + W)
    # the same sweep with the symmetry reduction switched off: every design point as one dense n x n problem
    Pg = min(P, args.general_points)
    ms_gen = None
    if args.general_path:
        def step_general(s):
            sl = (chunk(s)[:Pg]).to(device)
            return run_step(grids_res[sl], freq_res[sl], case, device, False)
        ms_gen = timed(step_general, K, W, s0=2 * (K + W))
    layers_per_step = P * world                                           # one patterned layer per design point
    value = layers_per_step * K / (ms_res * 1e-3)
    e2e_value = layers_per_step * K / (ms_e2e * 1e-3)

    out = None
    if rank == 0:
        n = 2 * (2 * args.order + 1) ** 2
        # ---- stage split and roofline of the eigen stage (untimed extra step on rank 0)
        from torcwa_b200 import _lib
        # the eig stage of the HEADLINE step as it ran (symmetry blocks: 4 P matrices of ~n/4 when both mirrors are found); taken
        # BEFORE the CUPTI pass below -- an attached profiler doubles the host API time the QR loop of many small matrices is sensitive to
        live = eig_calls_of(lambda: step_resident(1))
        live_bytes = sum(nb_ * 16.0 * (n_ ** 3 / 3.0 + 2.0 * n_ * n_) for nb_, n_, _ in live)
        live_ms = sum(ms_ for _, _, ms_ in live)
        launches, per_kernel = count_my_launches(lambda: step_resident(0))
        sym_probe = None
        try:            # which symmetry the package found for this workload (one untimed small solve)
            import torcwa_b200
            sl0_ = chunk(0)[:8].to(device)
            sim_ = torcwa_b200.rcwa(freq=freq_res[sl0_], order=case["order"], L=case["L"], dtype=torch.complex64, device=device)
            sim_.add_input_layer(eps=case["eps_in"]); sim_.set_incident_angle(inc_ang=0.0, azi_ang=0.0)
            sim_.add_layer(thickness=300.0, eps=grids_res[sl0_])
            if sim_._sym not in (None, False):
                sym_probe = {"mirrors": list(sim_._sym.gens), "block_sizes": [sim_._sym.sizes[c] for c in sim_._sym.chars]}
            del sim_
        except Exception as ex:
            sym_probe = {"error": str(ex)[:200]}
        torch.cuda.empty_cache()
        sl0 = chunk(0).to(device)
        Ps = P
        sim_stage, A_eig = stage_times(case, grids_res[sl0][:Ps].contiguous(), freq_res[sl0][:Ps].contiguous(), device)
        torch.cuda.empty_cache()
        mv = matvec_roofline(A_eig)
        tz = tensor_roofline(A_eig, digits=int(os.environ.get('RCWA_B200_GEMM_DIGITS', '7')))
        del A_eig
        torch.cuda.empty_cache()
        b_eig = 16.0 * (n ** 3 / 3.0 + 2.0 * n * n)                      # SURVEY.md 8d, s = 16 (fp64 internals)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)"
        dom = max(per_kernel.items(), key=lambda kv: kv[1][1])[0] if launches else None
        t_h = sim_stage["hessenberg_alone_ms"]
        achieved_h = Ps * b_eig / (t_h * 1e-3) / 1e9
        achieved_eig = Ps * b_eig / (sim_stage["eig_ms"] * 1e-3) / 1e9
        traffic = None
        try:        # dram__bytes_read + dram__bytes_write of one captured launch (ncu --set full), as a ratio to its algorithmic bytes
            tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))["hb_matvec_kernel"]
            traffic = tr["dram_bytes_per_algorithmic_byte"] * mv["bytes_per_launch"]
        except Exception:
            pass
        cupti_mv = per_kernel.get("hb_matvec")
        # roofline of the eig stage = B_eig / t_eig over the WHOLE stage (Hessenberg + QR + eigenvectors);
        # the phases that are not HBM-bound (QR: fp64 tensor pipe + serial chain) pull it far below the streaming kernel's own
        # fraction, which is reported beside it.
        general = {"kernel": "rcwa_eig on the dense n = %d problems (symmetry_reduction=False), batch %d" % (n, Ps),
                   "achieved": achieved_eig, "frac": achieved_eig / hbm_peak, "ms_per_batch": sim_stage["eig_ms"], "algorithmic_bytes_per_matrix": b_eig}
        if sym_probe and "block_sizes" in sym_probe and live_ms > 0:
            head = {"kernel": "rcwa_eig (whole stage: blocked Hessenberg reduction + multishift QR/AED + eigenvectors) as the headline step runs it: "
                              "%s; algorithmic bytes B_eig = 16 (m^3/3 + 2 m^2) per m x m matrix (SURVEY.md 8d, s = 16: fp64 internals)"
                              % ", ".join("%d matrices of %d" % (nb_, n_) for nb_, n_, _ in live),
                    "achieved": live_bytes / (live_ms * 1e-3) / 1e9, "ms_per_batch": live_ms,
                    "note": "the symmetry reduction removes 15/16 of the dense path's bytes (general_path below: the same stage on the n x n problems); "
                            "what is left is 512 small problems whose QR phase is bound by the serial bulge-chase / deflation chains and the "
                            "shared-memory footprint of the pass kernel, not by HBM"}
        else:
            head = dict(general)
        roof = {"bound": "hbm", "kernel": head["kernel"],
                "achieved": head["achieved"], "peak": hbm_peak, "unit": "GB/s", "frac": head["achieved"] / hbm_peak,
                "traffic": None, "peak_source": peak_src, "ms_per_batch": head["ms_per_batch"], "note": head.get("note"),
                "general_path": general,
                "hessenberg_phase_whole": {"achieved": achieved_h, "frac": achieved_h / hbm_peak, "ms_per_batch": t_h, "path": "general (n = %d)" % n},
                "streaming_kernel_alone": {
                    "kernel": "hb_matvec_kernel: y = A[k0+1:n, j+1:n] u_j, one launch per column (%d sampled columns timed alone with CUDA events, batch %d)" % (mv["launches"], Ps),
                    "achieved": mv["gbs"], "frac": mv["gbs"] / hbm_peak, "bytes_per_launch": mv["bytes_per_launch"], "us_per_launch": mv["us_per_launch"],
                    "traffic_offline_ncu": traffic, "traffic_note": "dram bytes of ONE captured launch (profiles/roofline_traffic.json) scaled to this launch size; not measured in this run",
                    "in_step_cupti": None if not cupti_mv else {
                        "launches": cupti_mv[0], "us_per_launch": cupti_mv[1] / max(cupti_mv[0], 1),
                        "achieved": P * (b_eig - 32.0 * n * n) / max(cupti_mv[1], 1e-9) / 1e3}}}
        cpu = cpu_baseline(args.order, args.ref_dtype) if (args.cpu_baseline and world == 1) else None     # N = 1 only
        torch.cuda.empty_cache()
        cuda_ref = cuda_baseline(args.order) if (args.cuda_baseline and world == 1) else None
        out = {
            "metric": "layers/sec (order %dx%d, c64)" % (args.order, args.order), "value": value, "unit": "layers/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_res / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None,
            "dtype": ("f64 behind the complex64 API (results rounded to c64): eigensolver, solves and the dense products of the %s-size symmetry blocks on the fp64 "
                      "pipes (DMMA); the tcgen05 int8-digit GEMM (7 x 8 bit, fp64-grade) takes the dense products of the S-matrix stage only for K >= 768, "
                      "i.e. on the general path (general_path, roofline_tensor.tcgen05)" % "/".join(str(b) for b in sorted(set(sym_probe["block_sizes"]))))
                     if sym_probe and "block_sizes" in sym_probe else
                     "f64 eigensolver and solves; the S-matrix stage's dense products on tcgen05 int8 digits (7 x 8 bit, fp64-grade) behind the complex64 API; results rounded to c64",
            "data": "synthetic (Example1 cell, linear a-Si:H dispersion, 512 wavelengths 400-700 nm)",
            "config": {"workload": "BASELINE configs[1]: order %dx%d (n=%d), 1 patterned layer + SiO2 half space, 512-wavelength sweep, complex64 API"
                                   % (args.order, args.order, n), "points_per_step_per_gpu": P, "layers_per_point": 1,
                       "symmetry_reduction": sym_probe if sym_probe else "none found: general path",
                       "l2": "working set per step >> 126 MB L2; every step takes new wavelengths", "parallelism": "dp%d (sweep sharded, final all_gather)" % world},
            "e2e": {"value": e2e_value, "unit": "layers/s", "h2d_bytes_per_step": int(P * (300 * 300 * 8 + 4)), "d2h_bytes_per_step": int(P * world * 8),
                    "ms_per_step": ms_e2e / K},
            "gpu_launches": (launches * K) if launches else None,
            "gpu_launches_per_step": launches,
            "clocks": clk,
            "roofline": roof,
            "roofline_tensor": tz,
            "stage_ms_per_batch": dict(sim_stage, batch=Ps),
            "dominant_kernel_by_time": dom,
            "kernel_time_share": {k: round(v[1] / max(sum(x[1] for x in per_kernel.values()), 1e-9), 4) for k, v in
                                  sorted(per_kernel.items(), key=lambda kv: -kv[1][1])[:8]} if launches else per_kernel,
            "kernel_launches_and_avg_us": {k: [v[0], round(v[1] / max(v[0], 1), 1)] for k, v in
                                           sorted(per_kernel.items(), key=lambda kv: -kv[1][1])[:8]} if launches else None,
            "general_path": None if ms_gen is None else {
                "value": Pg * world * K / (ms_gen * 1e-3), "unit": "layers/s", "ms_per_step": ms_gen / K, "points_per_step_per_gpu": Pg,
                "note": "the same sweep with symmetry_reduction=False: each design point one dense n x n problem (the path roofline.general_path, "
                        "roofline_tensor and stage_ms_per_batch describe; kernel_time_share and gpu_launches describe the headline step)"},
            "cpu_baseline": cpu,
            "cuda_baseline": cuda_ref,
            "vs_reference_cuda": None if not cuda_ref or "c64" not in cuda_ref or "layers_per_s" not in cuda_ref["c64"] else {
                "value_over_c64": value / world / cuda_ref["c64"]["layers_per_s"],
                "value_over_c128": None if "layers_per_s" not in cuda_ref.get("c128", {}) else value / world / cuda_ref["c128"]["layers_per_s"],
                "note": "per-GPU throughput of this arm / the unmodified reference on the same GPU in the same run"},
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def eig_calls_of(fn):
    """(batch, n, ms) of every rcwa_eig call one step makes, timed with CUDA events on the stream the step runs on."""
    from torcwa_b200 import _lib
    real, calls = _lib.eig, []

    def timed_eig(A, after_reduction=None):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = real(A, after_reduction)
        e1.record()
        calls.append((int(A.shape[0]), int(A.shape[1]), e0, e1))
        return out
    _lib.eig = timed_eig
    try:
        fn()
        torch.cuda.synchronize()
    finally:
        _lib.eig = real
    return [(nb, n, a.elapsed_time(b)) for nb, n, a, b in calls]


def stage_times(case, grids, freq, device):
    """CUDA-event timing of the stages of one batch through the C ABI wrappers (rank 0, untimed region)."""
    import torcwa_b200
    from torcwa_b200 import _lib
    P = grids.shape[0]
    sim = torcwa_b200.rcwa(freq=freq, order=case["order"], L=case["L"], dtype=torch.complex64, device=device)
    sim.add_input_layer(eps=case["eps_in"])
    sim.set_incident_angle(0.0, 0.0)
    o = case["order"][0]

    def ev():
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e
    res = {}
    A = None
    for rep in range(2):
        e = [ev()]
        E = _lib.convmat(grids, o, o, nb=P); e.append(ev())
        eta, _i = _lib.inverse(E); e.append(ev())
        Pm, Q = _lib.pq_assemble(eta, E, sim._kx, sim._ky, mu_scalar=torch.ones(P, dtype=torch.complex128, device=device))
        A = _lib.zgemm(Pm, Q); e.append(ev())
        del E, eta, Pm
        H = A.clone(); e.append(ev())
        _lib.hessenberg_(H); e.append(ev())
        del H
        lam, Wv, info = _lib.eig(A); e.append(ev())
        kz = _lib.kz_branch(lam)
        om = sim._omega64.expand(P).contiguous()
        th = torch.full((P,), 300.0, dtype=torch.float64, device=device)
        S11, S21, _i = _lib.layer_smatrix(Wv, kz, Q, sim._Vf_inv, om, th); e.append(ev())
        del Wv, Q
        S, _i = _lib.redheffer_bdleft(sim._Sin, [S11, S21, S21, S11]); e.append(ev())
        torch.cuda.synchronize()
        names = ["convmat_ms", "inv_eps_ms", "pq_and_product_ms", "_clone", "hessenberg_alone_ms", "eig_ms", "layer_smatrix_ms", "redheffer_ms"]
        res = {names[i]: e[i].elapsed_time(e[i + 1]) for i in range(len(names)) if not names[i].startswith("_")}
        st = _lib.last_eig_stats.cpu().numpy()
        res["qr_sweeps_per_matrix"] = float(st[:, 0].mean())
        res["qr_passes_max"] = int(st[:, 1].max())
        res["eig_info_max"] = int(info.abs().max())
        del S, S11, S21, lam, kz
        if rep == 0:
            del A
    pr = _lib.last_eig_profile.cpu().numpy().astype(float)
    seg = ["sweep_start", "chase_window", "small_block_slice", "aed_schur_slice", "aed_scan_slice", "aed_finish"]
    res["qr_pass_segments_mean_per_matrix"] = {seg[k]: {"launches": float(pr[:, k, 0].mean()), "sm_cycles": float(pr[:, k, 1].mean())} for k in range(6)}
    return res, A      # A: scratch contents after rcwa_eig (only its size matters to the roofline probes)


def matvec_roofline(A):
    """Kernel-only timing of the HBM-bound kernel: launches of hb_matvec_kernel for sampled columns, alone on the stream,
    CUDA events around them.  Algorithmic bytes of a launch = 16 (n-k0-1)(n-j-1) nb (what the kernel reads once)."""
    from torcwa_b200 import _lib
    nb, n = A.shape[0], A.shape[1]
    ws = _lib.eig_workspace(n, nb, A.device)
    ws.zero_()
    hb = _lib.load().rcwa_hessenberg_panel_width()
    cols = list(range(0, n - 2, 61))
    for j in cols[:2]:
        _lib.matvec_probe(A, ws, j)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for j in cols:
        _lib.matvec_probe(A, ws, j)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    byts = sum(16.0 * (n - (j // hb) * hb - 1) * (n - j - 1) * nb for j in cols)
    return {"gbs": byts / (ms * 1e-3) / 1e9, "launches": len(cols), "bytes_per_launch": byts / len(cols), "us_per_launch": 1e3 * ms / len(cols)}


def tensor_roofline(A, digits=7):
    """Tensor pipes, measured in this run with CUDA events on batched n x n x n complex products:
      * fp64 (DMMA, mma.sync): our kernel, against the fp64 tensor peak MEASURED here as a sustained cuBLAS dgemm (8192^3,
        back to back for ~1 s) -- MEASURED_PEAKS.json has no fp64 entry;
      * tcgen05 (kind::i8, the engine of the complex64 API's S-matrix stage): the digit GEMM at `digits` and at 8 digits,
        int8 ops actually issued / time, against 2 x the measured dense bf16 rate of MEASURED_PEAKS.json (kind::i8 runs at
        twice the kind::f16 rate on sm_100a; there is no measured int8 entry)."""
    from torcwa_b200 import _lib
    nb = min(A.shape[0], 8)
    n = A.shape[1]
    X, Y = A[:nb].contiguous(), A[nb:2 * nb].contiguous() if A.shape[0] >= 2 * nb else A[:nb].clone()
    out = torch.empty_like(X)

    def t(fn, reps=3):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    ms = t(lambda: _lib.zgemm(X, Y, out=out))
    ms_c = t(lambda: torch.matmul(X, Y))
    fl = 8.0 * n ** 3 * nb
    # measured fp64 tensor peak: sustained cuBLAS dgemm
    D1 = torch.randn(8192, 8192, dtype=torch.float64, device=A.device)
    D2 = torch.randn(8192, 8192, dtype=torch.float64, device=A.device)
    ms_d = t(lambda: torch.matmul(D1, D2), reps=20)
    fp64_peak = 2.0 * 8192.0 ** 3 / ms_d / 1e9
    del D1, D2
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    bf16 = float(peaks.get("bf16_tflops", 1590.0))
    res = {"bound": "tensor", "kernel": "zgemm_grouped_kernel<32,128,+3M> (batched %d x n^3 complex product, n=%d; fp64 DMMA mma.sync, tcgen05 has no f64 kind)" % (nb, n),
           "achieved": fl / ms / 1e9, "peak": fp64_peak, "unit": "TFLOP/s", "frac": fl / ms / 1e9 / fp64_peak,
           "peak_source": "measured in this run: sustained cuBLAS dgemm 8192^3 (20 back to back)", "cublas_zgemm_same_shape": fl / ms_c / 1e9,
           "note": "flops counted as 8 n^3 per complex product; the 3-multiplication kernel issues 6 n^3"}
    tc = {}
    for dg in sorted({int(digits), 8} - {0}):
        try:
            ms_t = t(lambda: _lib.zgemm_tc(X, Y, slices=dg, out=out))
            ops = 2.0 * 3 * (dg * (dg + 1) // 2) * n ** 3 * nb
            tc["digits_%d" % dg] = {"ms": ms_t, "speedup_vs_dmma": ms / ms_t, "equivalent_fp64_tflops": fl / ms_t / 1e9,
                                    "int8_tops_issued": ops / ms_t / 1e9, "frac_of_int8_peak": ops / ms_t / 1e9 / (2.0 * bf16)}
        except Exception as ex:
            tc["digits_%d" % dg] = {"error": str(ex)[:200]}
    res["tcgen05"] = {"kernel": "tc_gemm_kernel (tcgen05.mma kind::i8, TMEM accumulators, TMA-staged digit planes) incl. the digit split of both operands",
                      "int8_peak_tops": 2.0 * bf16, "int8_peak_source": "2 x bf16_tflops of MEASURED_PEAKS.json" if peaks else "2 x fallback 1590",
                      "tensor_pipe_active_pct_ncu": "see profiles/ (ncu sm__pipe_tensor_cycles_active of tc_gemm_kernel)", **tc}
    return res


# ------------------------------------------------------------------------------------------- CPU legs
def oracle_point(order, cdtype, lam=532.0):
    """One design point (one patterned layer) through the oracle = the reference's dense CPU algebra."""
    from oracle import cases as C
    from oracle.rcwa_oracle import OracleSim
    case = dict(C.CASES["ex1_o15"])
    case["order"] = [order, order]
    case["lam"] = lam
    t0 = time.perf_counter()
    sim = C.run_case(lambda **kw: OracleSim(**kw), case, cdtype)
    t = sim.S_parameters([0, 0])
    return time.perf_counter() - t0, complex(t[0])


REF_DTYPE_NOTE = ("reference CPU arithmetic timed in complex128: its complex64 path hits an MKL cgetri/cgetrf slow path on the 4N x 4N "
                  "coupling matrix (51 s/point on 8 cores in the build container, 455 s/point on this pool's 128-thread host, both measured), "
                  "so complex128 (17 s/point on 8 cores) is the FASTER, i.e. more favourable, reference")


def ref_threads():
    """MKL/LAPACK at this size (4N = 3844) gets SLOWER with very many threads: on this pool's 128-thread
    host the same design point took 358.8 s with 128 threads (measured 2026-09-25) against 17 s with 8
    threads in the build container.  The CPU legs therefore use at most 16 threads -- the setting that
    favours the reference -- and say so in their JSON."""
    return min(os.cpu_count() or 1, 16)


def cpu_point(order, cd, lam):
    """One design point on the host: the unmodified reference (baseline/_ref) if present, else the oracle port."""
    torcwa = reference_module()
    if torcwa is not None:
        dt, t, _ = reference_point(torcwa, order, cd, lam, torch.device("cpu"))
        return dt, t, "reference"
    dt, t = oracle_point(order, cd, lam)
    return dt, t, "port"


def cpu_baseline(order, ref_dtype="c128"):
    torch.set_num_threads(ref_threads())
    cd = torch.complex128 if ref_dtype == "c128" else torch.complex64
    dt, _, kind = cpu_point(order, cd, 532.0)
    return {"value": 1.0 / dt, "unit": "layers/s", "cores": torch.get_num_threads(), "host_cores": os.cpu_count(), "kind": kind,
            "sample": "1 design point (1 patterned layer, order %dx%d, %s) through %s, "
                      "%d threads (of %d; 128 threads measured 20x slower), %.1f s. %s"
                      % (order, order, ref_dtype, "the unmodified reference (baseline/_ref, device='cpu')" if kind == "reference" else "oracle/rcwa_oracle.py",
                         torch.get_num_threads(), os.cpu_count(), dt, REF_DTYPE_NOTE if ref_dtype == "c128" else "")}


def bench_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config != 2:
        return bench_reference_config(args)
    torch.set_num_threads(ref_threads())
    K, W = args.steps, args.warmup
    budget = args.ref_budget_s
    t_start = time.perf_counter()
    lams = np.linspace(LAM0, LAM1, N_SWEEP)
    cd = torch.complex128 if args.ref_dtype == "c128" else torch.complex64
    kind, warm_done = "port", 0
    for s in range(W):                     # warm-ups as requested, as long as the time budget allows two timed steps after them
        dt, _, kind = cpu_point(args.order, cd, float(lams[s]))
        warm_done += 1
        if time.perf_counter() - t_start + 3 * dt > budget:
            break
    done, total = 0, 0.0
    for s in range(K):
        dt, _, kind = cpu_point(args.order, cd, float(lams[(W + s) % N_SWEEP]))
        total += dt
        done += 1
        if time.perf_counter() - t_start + dt > budget:
            break
    value = done / total
    n = 2 * (2 * args.order + 1) ** 2
    impl = "the unmodified reference (baseline/_ref: torcwa.rcwa, device='cpu')" if kind == "reference" else "oracle/rcwa_oracle.py (port)"
    print(json.dumps({
        "impl": "reference", "metric": "layers/sec (order %dx%d, c64)" % (args.order, args.order), "value": value, "unit": "layers/s",
        "n_gpus": args.gpus, "steps": done, "requested_steps": K, "warmup": warm_done, "requested_warmup": W, "ms_per_step": 1e3 * total / done, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.ref_dtype, "data": "synthetic (same cell and sweep as the B200 arm)",
        "note": REF_DTYPE_NOTE if args.ref_dtype == "c128" else "",
        "config": {"workload": "BASELINE configs[1] unit: order %dx%d (n=%d), 1 patterned layer + SiO2 half space, one wavelength per step, through %s"
                               % (args.order, args.order, n, impl), "points_per_step": 1},
        "cpu_baseline": {"value": value, "unit": "layers/s", "cores": torch.get_num_threads(), "host_cores": os.cpu_count(), "kind": kind,
                         "sample": "%d x 1 design point through %s on %d torch threads (steps and warm-ups capped by --ref-budget-s %.0f)"
                                   % (done, impl, torch.get_num_threads(), budget)},
        "e2e": {"value": value, "unit": "layers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------- BASELINE configs[2] and configs[4]
SU8 = 1.6 ** 2


def config3_inputs(P):
    """Example1-1 style stack (Example1-1.ipynb:58-69,159-177): four a-Si:H bars (180 x 100 nm) in SU-8 rotated by
    0 / 30 / 60 / 90 degrees, 200 nm each, separated by 100 nm SU-8 spacers = 8 layers; wavelengths 600-700 nm."""
    import math
    import torcwa_b200
    geo = torcwa_b200.geometry(Lx=300.0, Ly=300.0, nx=300, ny=300, edge_sharpness=1000.0, dtype=torch.float64, device=torch.device("cpu"))
    masks = torch.stack([geo.rectangle(Wx=180.0, Wy=100.0, Cx=150.0, Cy=150.0, theta=th) for th in (0.0, math.pi / 6, math.pi / 3, math.pi / 2)])
    lams = torch.linspace(600.0, 700.0, N_SWEEP, dtype=torch.float64)
    return masks, lams


def step_config3(masks_dev, lam_dev, order, device):
    import torcwa_b200
    P = lam_dev.shape[0]
    e = torch.tensor([eps_si(float(l)) for l in lam_dev.cpu()], dtype=torch.complex128, device=device)
    sim = torcwa_b200.rcwa(freq=1.0 / lam_dev, order=[order, order], L=[300.0, 300.0], dtype=torch.complex128, device=device)
    sim.add_input_layer(eps=1.46 ** 2)
    sim.set_incident_angle(inc_ang=0.0, azi_ang=0.0)
    for k in range(4):
        grid = masks_dev[k][None] * e[:, None, None] + (1.0 - masks_dev[k])[None] * SU8
        sim.add_layer(thickness=200.0, eps=grid)
        del grid
        sim.add_layer(thickness=100.0, eps=SU8)
    sim.solve_global_smatrix()
    _LAST_SYMMETRY["config3"] = ({"group": list(sim._sym.gens), "block_sizes": [sim._sym.sizes[c] for c in sim._sym.chars]}
                                 if sim._sym not in (None, False) else "none found: general path")
    return sim.S_parameters(orders=[0, 0], direction="forward", port="transmission", polarization="xx", ref_order=[0, 0])


_LAST_SYMMETRY = {}


def config5_density(P, nx, ny, seed0=333):
    """Blurred uniform noise in (0, 1), one per geometry, CPU generator for portability (Example6.ipynb:926-930 uses the
    device RNG)."""
    out = []
    fx = torch.fft.fftfreq(nx, dtype=torch.float32)[:, None]
    fy = torch.fft.fftfreq(ny, dtype=torch.float32)[None, :]
    blur = torch.exp(-((fx * 40.0) ** 2 + (fy * 40.0) ** 2))
    for b in range(P):
        g = torch.Generator().manual_seed(seed0 + b)
        r = torch.rand(nx, ny, generator=g, dtype=torch.float32)
        r = 0.5 * (r + torch.flip(r, dims=[1]))
        sm = torch.fft.ifft2(torch.fft.fft2(r) * blur).real
        sm = (sm - sm.min()) / (sm.max() - sm.min())
        out.append(0.1 + 0.8 * sm)
    return torch.stack(out)


def step_config5(rho_dev, order, device):
    """Example6's iteration (Example6.ipynb:67-82,938-945): density -> permittivity -> one 300 nm layer on glass ->
    FoM = sum |t(1,0)|^2 over xx, yy, xy, yx -> backward; returns (FoM per geometry summed, gradient w.r.t. the densities)."""
    import torcwa_b200
    P = rho_dev.shape[0]
    rho = rho_dev.clone().requires_grad_(True)
    lam = torch.full((P,), 532.0, dtype=torch.float32, device=device)
    sim = torcwa_b200.rcwa(freq=1.0 / lam, order=list(order), L=[700.0, 300.0], dtype=torch.complex64, device=device)
    sim.add_input_layer(eps=1.46 ** 2)
    sim.set_incident_angle(inc_ang=0.0, azi_ang=0.0)
    eps = rho.to(torch.complex64) * complex(SI_EPS[532.0]) + (1.0 - rho)
    sim.add_layer(thickness=300.0, eps=eps)
    sim.solve_global_smatrix()
    fom = 0.0
    for pol in ("xx", "yy", "xy", "yx"):
        t = sim.S_parameters(orders=[1, 0], direction="forward", port="transmission", polarization=pol, ref_order=[0, 0])
        fom = fom + (t.abs() ** 2).sum()
    fom.backward()
    return fom.detach(), rho.grad


def bench_config(args):
    """--config 3 / 5: the same JSON schema as the headline config, with a simpler extra-metrics section."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    P = args.points
    if args.config == 3:
        order = args.order
        masks, lams = config3_inputs(P)
        masks_host, lams_host = masks.pin_memory(), lams.pin_memory()
        masks_res, lams_res = masks.to(device), lams.to(device)

        def chunk(s):
            return (torch.arange(P) + (s * world + rank) * P) % N_SWEEP

        def step_res(s):
            return step_config3(masks_res, lams_res[chunk(s).to(device)], order, device)

        def step_e2e(s):
            m = masks_host.to(device, non_blocking=True)
            l = lams_host[chunk(s)].pin_memory().to(device, non_blocking=True)
            return step_config3(m, l, order, device)
        layers_per_point, h2d = 8, int(masks.numel() * 8 + P * 8)
        n = 2 * (2 * order + 1) ** 2
        metric = "layers/sec (order %dx%d, 8 stacked layers, c128)" % (order, order)
        workload = ("BASELINE configs[2]: order %dx%d (n=%d), 8 layers per point (4 rotated a-Si:H bars in SU-8 + 4 homogeneous spacers, "
                    "Example1-1 style), complex128, %d wavelengths per step per GPU" % (order, order, n, P))
        dtype = "f64 (complex128 API: every stage on the fp64 pipes)"
    else:
        order = (args.order, args.order) if args.order_y is None else (args.order, args.order_y)
        nx, ny = 700, 300
        rho = config5_density(P * 4, nx, ny)
        rho_host, rho_res = rho.pin_memory(), rho.to(device)

        def chunk(s):
            return (torch.arange(P) + (s * world + rank) * P) % rho.shape[0]

        def step_res(s):
            return step_config5(rho_res[chunk(s).to(device)], order, device)[0]

        def step_e2e(s):
            r = rho_host[chunk(s)].pin_memory().to(device, non_blocking=True)
            f, g = step_config5(r, order, device)
            return torch.stack([f, g.abs().sum()])
        layers_per_point, h2d = 1, int(P * nx * ny * 4)
        n = 2 * (2 * order[0] + 1) * (2 * order[1] + 1)
        metric = "layers/sec forward+backward (order %dx%d, c64)" % order
        workload = ("BASELINE configs[4]: order %dx%d (n=%d), topology-optimisation step (Example6): density -> layer -> FoM = sum |t(1,0)|^2 "
                    "-> autograd backward to the 700 x 300 density, complex64 API, %d geometries per step per GPU" % (order[0], order[1], n, P))
        dtype = "f64 internals behind the complex64 API (differentiable pipeline: autograd over the C-ABI primitives)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K, W, s0=0, fetch=False):
        for s in range(W):
            r = fn(s0 + s)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(K):
            r = fn(s0 + W + s)
            if fetch:
                r = r.cpu()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms
    K, W = args.steps, args.warmup
    clocks = ClockSampler(local)
    clocks.start()
    ms_res = timed(step_res, K, W)
    clk = clocks.stop()
    ms_e2e = timed(step_e2e, K, 1, s0=K + W, fetch=True)
    if rank == 0:
        live = eig_calls_of(lambda: step_res(1))          # the eig calls of one step as they ran (before the CUPTI pass, see bench_b200)
        launches, per_kernel = count_my_launches(lambda: step_res(0))
        value = P * world * layers_per_point * K / (ms_res * 1e-3)
        tot_k = max(sum(v[1] for v in per_kernel.values()), 1e-9) if launches else 1.0
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        mvk = per_kernel.get("hb_matvec") if launches else None
        stage_bytes = sum(nb_ * 16.0 * (m_ ** 3 / 3.0 + 2.0 * m_ * m_) for nb_, m_, _ in live)
        stage_ms = sum(ms_ for _, _, ms_ in live)
        mv_bytes = sum(nb_ * 16.0 * m_ ** 3 / 3.0 for nb_, m_, _ in live)
        roof = None if not live or stage_ms <= 0 else {
            "bound": "hbm", "kernel": "rcwa_eig (whole stage) as the step runs it: %s; algorithmic bytes 16 (m^3/3 + 2 m^2) per m x m matrix (SURVEY.md 8d)"
                                      % ", ".join("%d x %d" % (nb_, m_) for nb_, m_, _ in live),
            "achieved": stage_bytes / (stage_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": stage_bytes / (stage_ms * 1e-3) / 1e9 / hbm_peak,
            "traffic": None, "ms_per_step": stage_ms, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback (B200_PROFILING.md)",
            "streaming_kernel_in_step": None if not mvk else {
                "kernel": "hb_matvec_kernel (streaming mat-vec of the blocked Hessenberg reduction), CUPTI durations of all its launches in one step",
                "achieved": mv_bytes / mvk[1] / 1e3, "frac": mv_bytes / mvk[1] / 1e3 / hbm_peak, "algorithmic_bytes": "16 m^3 / 3 per matrix"}}
        cpu = None
        if args.cpu_baseline and world == 1:
            cpu = cpu_baseline_config(args)
        print(json.dumps({
            "metric": metric, "value": value, "unit": "layers/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_res / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype,
            "data": "synthetic (rotated bars with linear a-Si:H dispersion)" if args.config == 3 else "synthetic (seeded blurred-noise densities)",
            "config": {"workload": workload, "points_per_step_per_gpu": P, "layers_per_point": layers_per_point,
                       "symmetry_reduction": _LAST_SYMMETRY.get("config3") if args.config == 3 else "not used by the differentiable path",
                       "l2": "working set per step >> 126 MB L2", "parallelism": "dp%d" % world},
            "e2e": {"value": P * world * layers_per_point * K / (ms_e2e * 1e-3), "unit": "layers/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": int(P * 8 if args.config == 3 else 16), "ms_per_step": ms_e2e / K},
            "gpu_launches": (launches * K) if launches else None, "gpu_launches_per_step": launches, "clocks": clk, "roofline": roof,
            "kernel_time_share": {k: round(v[1] / tot_k, 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][1])[:8]} if launches else per_kernel,
            "cpu_baseline": cpu}))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline_config(args):
    """CPU leg of configs 3 / 5 (one layer of one design point, extrapolated)."""
    torch.set_num_threads(ref_threads())
    if args.config == 3:
        dt, _, kind = cpu_point(args.order, torch.complex128, 650.0)
        return {"value": 1.0 / dt, "unit": "layers/s", "cores": torch.get_num_threads(), "host_cores": os.cpu_count(), "kind": kind,
                "sample": "ONE patterned layer of ONE design point at order %dx%d, complex128 (%.1f s), taken as the per-layer cost of the 8-layer stack "
                          "(BASELINE.md 3.1 allows the extrapolation; the homogeneous layers cost the reference the same dense algebra)" % (args.order, args.order, dt)}
    torcwa = reference_module()
    if torcwa is None:
        return {"unavailable": "baseline/_ref is missing (the oracle port has no autograd leg)"}
    ox, oy = 15, 8                                                   # Example6's own order; order 25x25 takes > 15 min per forward on the host
    rho = config5_density(1, 700, 300)[0].requires_grad_(True)
    t0 = time.perf_counter()
    sim = torcwa.rcwa(freq=1 / torch.tensor(532.0), order=[ox, oy], L=[700.0, 300.0], dtype=torch.complex64, device=torch.device("cpu"))
    sim.add_input_layer(eps=1.46 ** 2)
    sim.set_incident_angle(inc_ang=0.0, azi_ang=0.0)
    sim.add_layer(thickness=300.0, eps=rho.to(torch.complex64) * complex(SI_EPS[532.0]) + (1.0 - rho))
    sim.solve_global_smatrix()
    fom = 0.0
    for pol in ("xx", "yy", "xy", "yx"):
        fom = fom + (sim.S_parameters(orders=[1, 0], direction="forward", port="transmission", polarization=pol, ref_order=[0, 0]).abs() ** 2).sum()
    fom.backward()
    dt = time.perf_counter() - t0
    n_ref = 2 * (2 * ox + 1) * (2 * oy + 1)
    oyy = args.order if args.order_y is None else args.order_y
    n_here = 2 * (2 * args.order + 1) * (2 * oyy + 1)
    scale = (n_here / n_ref) ** 3
    return {"value": 1.0 / (dt * scale), "unit": "layers/s", "cores": torch.get_num_threads(), "host_cores": os.cpu_count(), "kind": "reference",
            "sample": "ONE forward+backward of the unmodified reference at Example6's own order [15,8] (n=%d, complex64, %.1f s), scaled by (n/n_ref)^3 = %.1f "
                      "to this order (n=%d): an EXTRAPOLATION (BASELINE.md 3.1), the full size takes > 15 min per forward on the host" % (n_ref, dt, scale, n_here)}


def bench_reference_config(args):
    c = cpu_baseline_config(args)
    if "value" not in c:
        print(json.dumps({"impl": "reference", "unavailable": c.get("unavailable", "?")}))
        return
    print(json.dumps({"impl": "reference", "metric": "layers/sec (config %d)" % args.config, "value": c["value"], "unit": "layers/s", "n_gpus": args.gpus,
                      "steps": 1, "warmup": 0, "ms_per_step": 1e3 / c["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "c128" if args.config == 3 else "c64", "data": "synthetic", "config": {"workload": c["sample"]}, "cpu_baseline": c,
                      "e2e": {"value": c["value"], "unit": "layers/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 5], help="BASELINE.json configs[config-1]: 2 = headline (order 15, c64 sweep), "
                    "3 = order 21, 8 stacked layers, complex128, 5 = forward + autograd backward (Example6)")
    ap.add_argument("--points", type=int, default=None, help="design points per step per GPU (default 128 / 16 / 4 for config 2 / 3 / 5)")
    ap.add_argument("--order", type=int, default=None, help="Fourier order (default 15 / 21 / 25 for config 2 / 3 / 5)")
    ap.add_argument("--order-y", type=int, default=None, help="config 5 only: second order (Example6 uses [15, 8])")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-general-path", dest="general_path", action="store_false", help="skip the extra timing of the sweep with the symmetry reduction off")
    ap.add_argument("--general-points", type=int, default=128, help="design points per step of the general-path leg (its footprint is ~0.7 GB per point)")
    ap.add_argument("--no-cuda-baseline", dest="cuda_baseline", action="store_false", help="skip the reference's PyTorch-CUDA path (unmodified baseline/_ref on this GPU)")
    ap.add_argument("--ref-budget-s", type=float, default=240.0)
    ap.add_argument("--ref-dtype", default="c128", choices=["c64", "c128"], help="arithmetic of the CPU reference legs (see REF_DTYPE_NOTE)")
    args = ap.parse_args()
    if args.points is None:
        args.points = {2: 128, 3: 16, 5: 4}[args.config]
    if args.order is None:
        args.order = {2: 15, 3: 21, 5: 25}[args.config]
    if args.impl == "reference":
        bench_reference(args)
    elif args.config != 2:
        bench_config(args)
    else:
        bench_b200(args)


if __name__ == "__main__":
    main()
