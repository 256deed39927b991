#!/usr/bin/env python3
"""bench.py -- SK pairs/s for the H0/S build + force contraction (BASELINE.json metric) and the
SCC-step times, on N B200s of one node.

A step = one pass of the pair hot path over one synthetic geometry of BASELINE config 4 (fec4096):
device-side pair schedule, buildH0 + buildS, derivative_shift (SCC form, collinear spin) for every
neighbour pair.
  value  device-resident: inputs in HBM, CUDA events on the launch stream, L2 flushed between steps, max over ranks;
  e2e    the same step with HOST (pinned) inputs and outputs, copies inside the timed region.
Extra objects: roofline (pair build, HBM bound), roofline_k5 (force kernel, FP64), roofline_dm (K3, FP64
tensor pipe), kernels (device time per routine with the reference's timer names), scc_step (addShift ->
unpack -> [eigensolve] -> DM -> pack -> Mulliken, with the eigensolve at the full size from a child
process), kpoints_si512 (BASELINE config 3: the complex chain, (k,spin) round-robin over the GPUs),
md_water2046 (config 5: velocity-Verlet steps with list rebuild and SCC restart), multi_gpu (N > 1: the
reference's split pair build + all-reduce measured against the replicated build), cpu_baseline (the oracle
= restatement of the reference CPU path, rank 0, N = 1), neighbour_list, coulomb_ewald.

`--impl reference` times that CPU restatement itself (the reference is Fortran and cannot be built in
this image: no Fortran compiler) with all host threads, same config, same steps.

Multi-GPU (one process per GPU, torchrun): every GPU keeps ONE context.  The pair build is replicated
(every GPU needs all of H0 / S for its dense work; the split build + all-reduce is timed beside it in
multi_gpu), the force contraction takes every N-th chunk of the pair schedule
(dftbp_b200_set_pair_partition) with an NCCL all-reduce of deriv(3,nAtom), the (k-point, spin) dense work
goes round-robin with an NCCL all-reduce of the sparse rho, column panels of the density matrix beyond.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sk_pairs_per_s_h0s_plus_force"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="fec4096")
    ap.add_argument("--no-dense", action="store_true", help="skip the SCC-step (dense) section")
    ap.add_argument("--no-eigensolve", action="store_true", help="skip the full-size eigensolve (child process, ~40 s)")
    ap.add_argument("--no-n2", action="store_true", help="skip the neighbour-list report (tools/n2_probe.py)")
    ap.add_argument("--no-kpoints", action="store_true", help="skip the config-3 section (si512, 32 k-points)")
    ap.add_argument("--md", type=int, default=10, help="velocity-Verlet steps of the config-5 section (0: skip; 100 = BASELINE)")
    ap.add_argument("--pair-only", action="store_true", help="only the headline pair step (ncu sessions)")
    ap.add_argument("--flush-l2", action="store_true",
                    help="write a 256 MB buffer between timed steps (the step itself already streams 460 MB > L2; the "
                         "dirty lines of such a flush make the first kernel of the step pay for their write-back)")
    return ap.parse_args()


def config_for(workload_name, geo, n_spin, flush=False):
    """The `config` object of the JSON line; identical for the b200 and the reference arm."""
    ss = int(geo["sparseSize"])
    stream_mb = ss * 8 * (2 + n_spin + 1 + 2) / 1e6     # H0, S out; rho_s, Erho in; -wS out and in
    return {"workload": workload_name, "nAtom": int(geo["nAtom"]), "nOrb": int(geo["nOrb"]),
            "pairs": int(geo["nNeighbourSK"].sum()), "sparseSize": ss, "n_spin": int(n_spin),
            "sk_tables": "synthetic .skf, 0.02 bohr x 500",
            "step": "new coordinates on an unchanged neighbour list + buildH0 + buildS + derivative_shift(SCC, %d spins, "
                    "diagonal shift blocks); GPU arm: the device-side pair schedule is refreshed per step and sorted "
                    "again every 16th update" % n_spin,
            "l2": ("flushed between timed steps (256 MB memset)" if flush else
                   "no explicit flush: every step streams %.0f MB (H0 + S out, rho / Erho / -wS in and out) through the "
                   "126 MB L2, each array larger than what is left of it at its next use" % stream_mb)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The
    sampler is started before the warm-up (nvidia-smi needs a few hundred ms to come up, the timed
    region of this benchmark is only milliseconds long) and read after the per-routine timing loop;
    a reader thread keeps the lines, `mark()` brackets the part that belongs to the GPU-busy window."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines, self.t0, self.thread = index, None, [], None, None

    def start(self):
        import threading
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def reader():
            try:
                for line in self.proc.stdout:
                    self.lines.append((time.perf_counter(), line))
            except Exception:
                pass
        self.thread = threading.Thread(target=reader, daemon=True)
        self.thread.start()

    def mark(self):
        """start of the GPU-busy window"""
        self.t0 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        try:
            deadline = time.perf_counter() + 3.0       # at least one sample inside the window
            while time.perf_counter() < deadline and not any(t >= (self.t0 or 0.0) for t, _ in self.lines):
                time.sleep(0.02)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in list(self.lines):
            if self.t0 is not None and t < self.t0:
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons),
                "window": "timed steps + per-routine timing loop (GPU busy throughout)"}


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of a kernel, from the newest committed
    `ncu --set full` summary under profiles/ (tools/ncu_summary.py); None if there is none.  The
    capture is a separate run of tools/gpu_probe.py under ncu, never the timed run."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_*.txt")), key=os.path.getmtime)
    for path in reversed(files):
        txt = open(path).read()
        for block in txt.split("## launch")[1:]:
            if kernel_substr not in block.splitlines()[0]:
                continue
            rd = re.search(r"dram read\s+([0-9.]+)\s+(\w+)", block)
            wr = re.search(r"dram write\s+([0-9.]+)\s+(\w+)", block)
            if not (rd and wr):
                continue
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            total = float(rd.group(1)) * scale.get(rd.group(2), 1.0) + float(wr.group(1)) * scale.get(wr.group(2), 1.0)
            return {"bytes_per_launch": total, "source": os.path.relpath(path, ROOT)}
    return None


def algorithmic_bytes_k1(geo):
    """16*n1*n2 + 8 per pair (H0 + S blocks written, iNeighbour + iPair read) + 16*n1^2 per atom
    (on-site blocks); DESIGN.md section K1 / SURVEY.md 8(d)."""
    return 16.0 * geo["sparseSize"] + 8.0 * float(geo["nNeighbourSK"].sum())


def algorithmic_flops_k5(geo, sk, n_spin):
    """12 x (K1 flops per pair) + (6 + 2 nSpin (n1 + n2)) n1 n2 per pair, with
    K1 flops per pair and matrix = 16 nInteg (8-point interpolation) + 2 x (FMAs of the rotation
    sum_m V_|m| D2(b,m) D1(a,m) over the shell pairs), DESIGN.md section K5."""
    shells = sk.ang_shells
    sp = geo["species"]
    n_orb = [sum(2 * l + 1 for l in sh) for sh in shells]
    per = {}
    for i, s1 in enumerate(shells):
        for j, s2 in enumerate(shells):
            n_int = sum(min(a, b) + 1 for a in s1 for b in s2)
            rot = sum((2 * a + 1) * ((2 * min(a, b) + 1) + (2 * b + 1) * (2 * min(a, b) + 1)) for a in s1 for b in s2)
            k1 = 16.0 * n_int + 2.0 * rot
            per[(i, j)] = 12.0 * k1 + (6.0 + 2.0 * n_spin * (n_orb[i] + n_orb[j])) * n_orb[i] * n_orb[j]
    total = 0.0
    nn, nb, img = geo["nNeighbourSK"], geo["iNeighbour"], geo["img2CentCell"]
    for a in range(geo["nAtom"]):
        s1 = sp[a] - 1
        for k in range(1, nn[a] + 1):
            total += per[(s1, sp[img[nb[a, k] - 1] - 1] - 1)]
    return total


def cublas_dsyrk_ms(stream, n, k, a_ptr, lda, c_ptr, ldc, reps=2):
    """cuBLAS DSYRK (lower, C = A A^T, A n x k column-major) on the given stream, through ctypes: the library
    routine the reference's CPU path uses for the density matrix (dsyrk / zherk, densitymatrix.F90:566-584) --
    the comparison SURVEY's hard part 6 asks for.  Returns ms per call or None."""
    import ctypes
    import torch
    try:
        cb = ctypes.CDLL("libcublas.so.12")
    except OSError:
        try:
            cb = ctypes.CDLL("/usr/local/cuda/lib64/libcublas.so")
        except OSError:
            return None
    h = ctypes.c_void_p()
    if cb.cublasCreate_v2(ctypes.byref(h)) != 0:
        return None
    try:
        cb.cublasSetStream_v2(h, ctypes.c_void_p(stream.cuda_stream))
        one, zero = ctypes.c_double(1.0), ctypes.c_double(0.0)
        CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N = 0, 0

        def call():
            return cb.cublasDsyrk_v2(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, ctypes.c_int(n), ctypes.c_int(k),
                                     ctypes.byref(one), ctypes.c_void_p(a_ptr), ctypes.c_int(lda), ctypes.byref(zero),
                                     ctypes.c_void_p(c_ptr), ctypes.c_int(ldc))
        if call() != 0:
            return None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            call()
        e1.record(stream)
        e1.synchronize()
        return e0.elapsed_time(e1) / reps
    finally:
        cb.cublasDestroy_v2(h)


def scc_shift(rng, n_spin, n_atom, m_orb):
    """Block shift of a collinear-spin SCC run without DFTB+U / on-site corrections (BASELINE config
    4): diagonal mOrb x mOrb blocks, one value per orbital (hamiltonian.F90:148-293 -> shift.F90:232-309
    expand atom / shell potentials onto the block diagonal)."""
    sh = np.zeros((n_spin, n_atom, m_orb, m_orb))
    idx = np.arange(m_orb)
    sh[:, :, idx, idx] = 0.05 * rng.normal(size=(n_spin, n_atom, m_orb))
    return sh


def reference_arm(a, rank, world):
    """The reference's CPU path (oracle restatement, C + OpenMP, all host threads): same step, same
    config, --steps timed steps after --warmup warm-up steps (one step is about 0.2 s on 16 threads)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle, case
    oracle.set_num_threads()          # all host cores (torchrun exports OMP_NUM_THREADS=1)
    w, sk, geo, model, selfegy = case(a.workload)
    n_spin = w["n_spin"]
    rng = np.random.default_rng(1)
    s = oracle.build_s(model, geo)
    dm = s * rng.normal(size=(n_spin, geo["sparseSize"]))
    edm = s * rng.normal(size=geo["sparseSize"])
    shift = scc_shift(rng, n_spin, geo["nAtom"], model.mOrb)
    pairs = int(geo["nNeighbourSK"].sum())

    def step():
        oracle.build_h0(model, geo, selfegy)
        oracle.build_s(model, geo)
        oracle.derivative_shift(model, geo, dm, edm, shift=shift)
    for _ in range(a.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    dt = (time.perf_counter() - t0) / max(1, a.steps)
    val = pairs / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_for(a.workload, geo, n_spin),
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": oracle.num_threads(), "kind": "port",
                             "sample": "full workload, %d steps: buildH0 + buildS + derivative_shift(SCC, %d spins)"
                                       % (a.steps, n_spin)},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "restatement of the reference CPU path (OpenMP over atoms), not a DFTB+ build"}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------
# extra sections (each returns a dict for the JSON line; failures are reported, never fatal)
# ------------------------------------------------------------------------------------------------------

def section_kpoints(dev, stream, rank, world, dist, max_over_ranks, barrier):
    """BASELINE config 3: 512-atom Si, 4x4x4 Monkhorst-Pack (32 k-points), one SCC iteration of the complex
    chain per local (k,spin): unpackHS(H, k), unpackHS(S, k) -> Zhegvd -> density matrix -> packHS with the
    k-point weight; (k,spin) round-robin over the GPUs (type/parallelks.F90:78-84), all-reduce of the sparse
    rho (main.F90:4327), Mulliken."""
    import torch
    from dftbplus_b200 import workloads, schedule
    from dftbplus_b200.api import HotPath, new_dense
    w = workloads.BY_NAME["si512"]()
    sk, geo, _ = workloads.prepare(w)
    hp = HotPath(dev.index)
    hp.set_stream(stream.cuda_stream)
    sk.upload(hp)
    hp.set_geometry(geo)
    n, ss = geo["nOrb"], geo["sparseSize"]
    kpts, wts = w["kpoints"], w["kweights"]
    my = schedule.parallel_ks(world, rank, len(kpts), 1)
    f64 = dict(dtype=torch.float64, device=dev)
    n_occ = int(round(sum(sum(sk.occ[s - 1]) for s in w["species0"]) / 2))
    with torch.cuda.stream(stream):
        ham, over = torch.zeros(ss, **f64), torch.zeros(ss, **f64)
        hp.buildH0S(sk.selfegy, ham, over)
        H, S = new_dense(n, dtype=torch.complex128, device=dev), new_dense(n, dtype=torch.complex128, device=dev)
        D = new_dense(n, dtype=torch.complex128, device=dev)
        fill = torch.zeros(n, **f64)
        fill[:n_occ] = 2.0
        rho = torch.zeros(ss, **f64)
        qq = torch.zeros((geo["nAtom"], hp.mOrb), **f64)
        eig = torch.zeros(n, **f64)
    stream.synchronize()

    def iteration():
        with torch.cuda.stream(stream):
            rho.zero_()
            qq.zero_()
        for (iK, _iS) in my:
            kp, wk = kpts[iK - 1], float(wts[iK - 1])
            hp.unpackHS(ham, kpoint=kp, square=H)
            hp.unpackHS(over, kpoint=kp, square=S)
            hp.diagDenseMtx(H, S, eigen=eig)                  # H <- eigenvectors
            hp.getDensityMatrix(H, fill, dm=D)
            hp.packHS(rho, D, kpoint=kp, kweight=wk)
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(rho)
        hp.mulliken(over, rho, qq=qq)
    iteration()
    barrier()
    hp.timer_enable(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    e0.record(stream)
    for _ in range(reps):
        iteration()
    e1.record(stream)
    stream.synchronize()
    ms = max_over_ranks(e0.elapsed_time(e1) / reps)
    t = hp.timer_read()
    per = {k: (v[0] / v[1] if v[1] else None) for k, v in t.items() if v[1]}
    n_el = float(qq.sum().item())
    out = {"workload": "si512", "nAtom": int(geo["nAtom"]), "nOrb": int(n), "k_points": len(kpts),
           "ks_pairs_this_rank": len(my), "ms_per_scc_iteration": ms,
           "per_call_ms_rank0": {"unpackHS_cplx": per.get("sparseToDense"), "Zhegvd": per.get("diagonalization"),
                                 "densityMatrix_cplx": per.get("densityMatrix"), "packHS_cplx": per.get("denseToSparse")},
           "electrons_from_mulliken": n_el, "electrons_expected": 2.0 * n_occ,
           "note": "unpack(H,k), unpack(S,k), Zhegvd, DM, pack per local k-point; all-reduce rho; Mulliken"}
    del H, S, D
    hp.close()
    return out


def fermi_fillings_device(eig, n_elec, kT):
    """Fermi filling of one spin-degenerate channel on the device (the job of Efilling, dftb/etemp.F90:59-191):
    bisection of the chemical potential, 2 / (1 + exp((e - mu) / kT))."""
    import torch
    lo, hi = float(eig.min()) - 10.0 * kT, float(eig.max()) + 10.0 * kT
    for _ in range(60):
        mu = 0.5 * (lo + hi)
        n = float((2.0 / (1.0 + torch.exp(torch.clamp((eig - mu) / kT, -500.0, 500.0)))).sum())
        if n > n_elec:
            hi = mu
        else:
            lo = mu
    return 2.0 / (1.0 + torch.exp(torch.clamp((eig - 0.5 * (lo + hi)) / kT, -500.0, 500.0)))


def section_md(dev, stream, n_steps, potential=None):
    """BASELINE config 5 pattern: 2046-atom water box, velocity-Verlet steps; per step the device neighbour
    list (update + adopt), buildH0 + buildS, an SCC loop restarted from the previous charges (addShift, unpack
    H and S, Dsygvd, density matrix, pack, Mulliken), energy-weighted density matrix and the SCC force
    contraction.  The charge model of this harness is the on-site part of the SCC Hamiltonian only (shift_A =
    U_A dq_A on the block diagonal, linear mixing) -- the gamma / Ewald sums are the Fortran host's (SURVEY 8(f)
    N4); every kernel of the hot path runs as in a real step."""
    import torch
    from dftbplus_b200 import workloads
    from dftbplus_b200.api import HotPath, new_dense
    w = workloads.BY_NAME["water2046"]()
    sk, geo0, _ = workloads.prepare(w)
    hp = HotPath(dev.index)
    hp.set_stream(stream.cuda_stream)
    sk.upload(hp)
    cutoff = sk.cutoff()
    lat = w["latvec"]
    species0 = w["species0"]
    nAtom = len(species0)
    f64 = dict(dtype=torch.float64, device=dev)
    if potential is None:
        potential = os.environ.get("DFTBP_B200_MD_POTENTIAL", "onsite")
    list_cutoff = cutoff
    if potential == "reference":
        from dftbplus_b200 import coulomb as pc
        # one U per species (OrbitalResolvedSCC = No); Ewald parameters and gamma cutoffs: per-lattice host set-up
        hubm = pc.unique_hubbard([[workloads.SPECIES[nm][2][0]] * len(sk.ang_shells[i]) for i, nm in enumerate(w["names"])],
                                 [len(a) for a in sk.ang_shells])
        gcuts = pc.short_gamma_cutoffs(hubm)
        ew = pc.ewald_setup(lat)
        list_cutoff = max(cutoff, float(gcuts.max()))      # mCutOff: the list covers every cutoff
    rng = np.random.default_rng(3)
    mass = np.array([workloads.SPECIES[w["names"][s - 1]][4] for s in species0]) * 1822.888486
    hub = np.array([workloads.SPECIES[w["names"][s - 1]][2][0] for s in species0])
    q0 = np.array([sum(sk.occ[s - 1]) for s in species0], dtype=np.float64)
    n_occ = int(round(q0.sum() / 2))
    x = w["coord0"].copy()
    v = rng.normal(size=x.shape) * np.sqrt(0.00095 / mass)[:, None]       # ~300 K
    v -= (mass[:, None] * v).sum(axis=0) / mass.sum()
    dt = 0.5 * 41.341373 / 1.0                                               # 0.5 fs in atomic units
    mo = hp.mOrb
    n = int(geo0["nOrb"])
    with torch.cuda.stream(stream):
        H, S, D = new_dense(n, device=dev), new_dense(n, device=dev), new_dense(n, device=dev)
        fill = torch.zeros(n, **f64)
        eig = torch.zeros(n, **f64)
        deriv = torch.zeros((nAtom, 3), **f64)
        shift = torch.zeros((1, nAtom, mo, mo), **f64)
        hub_d = torch.from_numpy(hub).to(dev)
        q0_d = torch.from_numpy(q0).to(dev)
        didx = torch.arange(mo, device=dev)
        if potential == "reference":
            invr = new_dense(nAtom, device=dev)
            rlat_d, glat_d = torch.from_numpy(ew["rLat"]).to(dev), torch.from_numpy(ew["gLat"]).to(dev)
            dq_shell = torch.zeros((nAtom, hubm["mShell"]), **f64)
            sh_shell = torch.zeros((nAtom, hubm["mShell"]), **f64)
            sh_atom = torch.zeros(nAtom, **f64)
    dq = torch.zeros(nAtom, **f64)
    kT = 0.02        # the synthetic SK tables leave the box without a gap: Fermi smearing (hartree)
    scc_iters, t_list, t_build, t_scc, t_force, t_pot = [], 0.0, 0.0, 0.0, 0.0, 0.0
    mixer = None
    mixer_kind = os.environ.get("DFTBP_B200_MD_MIXER", "broyden")
    if mixer_kind == "broyden":
        hp.mixerInit(64)                 # modified Broyden with device-resident vectors (dftbp_b200_mixer_*)
        mixer = "device"
    elif mixer_kind == "broyden_host":
        from dftbplus_b200.mixer import BroydenMixer
        mixer = BroydenMixer(max_scc_iter=64)

    def forces(x):
        nonlocal dq, fill, t_list, t_build, t_scc, t_force, t_pot
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g = hp.updateNeighbourListAndSpecies(x, species0, list_cutoff, lat, sk_cutoff=cutoff, adopt=True, fetch=False)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if potential == "reference":
            # TCoulomb%updateCoords + TShortGamma%updateCoords on the device
            with torch.cuda.stream(stream):
                x_d = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
            hp.invRPeriodic(x_d, rlat_d, glat_d, ew["alpha"], ew["volume"], ew["maxR"], invr=invr)
            hp.shortGammaUpdate(hubm["nHubbU"], hubm["uniqHubbU"], gcuts)
            torch.cuda.synchronize()
            t_pot += time.perf_counter() - t1
            t1 = time.perf_counter()
        ss = g["sparseSize"]
        with torch.cuda.stream(stream):
            ham, over = torch.empty(ss, **f64), torch.empty(ss, **f64)
            hscc = torch.empty((1, ss), **f64)
            rho, erho = torch.zeros(ss, **f64), torch.zeros(ss, **f64)
            qq = torch.zeros((nAtom, mo), **f64)
        hp.buildH0S(sk.selfegy, ham, over)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        it = 0
        hist_x, hist_r = [], []
        if mixer == "device":
            hp.mixerReset(nAtom)         # the reference resets its mixer for every geometry
        elif mixer is not None:
            mixer.reset(nAtom)
        for it in range(1, 41):
            if potential == "reference":
                # TCoulomb%updateShifts + TShortGamma%updateShifts: V_A = (1/R dq)_A + shiftShell(1, A)
                with torch.cuda.stream(stream):
                    dq_shell.zero_()
                    dq_shell[:, 0] = dq
                hp.coulombShift(invr, dq, shift=sh_atom)
                hp.shortGammaShifts(hubm["iHubbU"], dq_shell, shift_shell=sh_shell)
                with torch.cuda.stream(stream):
                    shift.zero_()
                    shift[0, :, didx, didx] = (sh_atom + sh_shell[:, 0])[:, None]
            else:
                with torch.cuda.stream(stream):
                    shift.zero_()
                    shift[0, :, didx, didx] = (hub_d * dq)[:, None]
            hp.addShift(hscc, over, shift, is_input_zero=True)
            with torch.cuda.stream(stream):
                hscc[0].add_(ham)
            hp.unpackHS(hscc[0], square=H)
            hp.unpackHS(over, square=S)
            hp.diagDenseMtx(H, S, eigen=eig)
            # Efilling on the device (dftbp_b200_efilling): n_occ electrons over singly occupied levels, times
            # two for the spin-unpolarised run (getFillingsAndBandEnergies, main.F90:4597-4648)
            fill = hp.Efilling(eig, float(n_occ), kT)["filling"]
            with torch.cuda.stream(stream):
                fill.mul_(2.0)
            hp.getDensityMatrix(H, fill, dm=D)
            with torch.cuda.stream(stream):
                rho.zero_()
                qq.zero_()
            hp.packHS(rho, D)
            hp.mulliken(over, rho, qq=qq)
            with torch.cuda.stream(stream):
                dq_out = qq.sum(dim=1) - q0_d
                res = dq_out - dq
                err = float(res.abs().max().item())
                if os.environ.get("MD_DEBUG"):
                    print("scc it %d err %.3e  dq range %.3f %.3f  gap %.4f" % (it, err, float(dq_out.min()), float(dq_out.max()),
                          float(eig[n_occ] - eig[n_occ - 1])), flush=True)
                if err < 1e-5:
                    break
                if mixer == "device":
                    # modified Broyden mixing of the atomic charges, the reference's default; the vectors stay on the GPU
                    dq = dq.clone()
                    hp.mixerMix(dq, res.contiguous())
                    continue
                if mixer is not None:
                    dq = torch.from_numpy(mixer.mix(dq.cpu().numpy(), res.cpu().numpy())).to(dev)
                    continue
                # Anderson mixing over the last five (input, residual) pairs of this geometry
                hist_x.append(dq.clone()); hist_r.append(res.clone())
                if len(hist_x) > 5:
                    hist_x.pop(0); hist_r.pop(0)
                if len(hist_x) == 1:
                    dq = dq + 0.4 * res
                else:
                    dR = torch.stack([hist_r[-1] - r for r in hist_r[:-1]], dim=1)
                    dX = torch.stack([hist_x[-1] - xx for xx in hist_x[:-1]], dim=1)
                    gam = torch.linalg.lstsq(dR, res.unsqueeze(1)).solution.squeeze(1)
                    dq = dq - dX @ gam + 0.4 * (res - dR @ gam)
        scc_iters.append(it)
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        hp.getEDensityMatrix(H, fill, eig, dm=D)
        hp.packHS(erho, D)
        hp.derivative_shift(rho.unsqueeze(0), erho, shift=shift, deriv=deriv)
        stream.synchronize()             # deriv is written on the library's stream
        f = -deriv.cpu().numpy()
        t4 = time.perf_counter()
        t_list += t1 - t0; t_build += t2 - t1; t_scc += t3 - t2; t_force += t4 - t3
        return f

    f = forces(x)                    # step 0 (SCC from neutral atoms), not timed
    scc_iters.clear()
    t_list = t_build = t_scc = t_force = t_pot = 0.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_steps):
        v += 0.5 * dt * f / mass[:, None]
        x = x + dt * v
        f = forces(x)
        v += 0.5 * dt * f / mass[:, None]
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    out = {"workload": "water2046", "nAtom": int(nAtom), "nOrb": n, "steps": n_steps, "steps_per_s": n_steps / wall,
           "ms_per_step": wall / n_steps * 1e3, "scc_iterations_per_step": float(np.mean(scc_iters)),
           "scc_potential": ("reference: Ewald 1/R matrix + short-range gamma on the device (one U per species), list cutoff %.2f bohr"
                             % list_cutoff) if potential == "reference" else "on-site part only (U_A dq_A)",
           "ms_per_step_breakdown": {"neighbour_list_update_adopt": t_list / n_steps * 1e3,
                                     "ewald_matrix_and_gamma_values": t_pot / n_steps * 1e3,
                                     "buildH0S": t_build / n_steps * 1e3, "scc_loop": t_scc / n_steps * 1e3,
                                     "edm_pack_forces": t_force / n_steps * 1e3},
           "momentum_drift": float(np.abs((mass[:, None] * v).sum(axis=0)).max()),
           "timing": "host wall clock (the step contains cuSOLVER calls and small D2H reads)",
           "note": ("SCC potential of the reference on the device (dftbp_b200_invr_periodic, _short_gamma_update, _coulomb_shift, "
                    "_short_gamma_shifts), " if potential == "reference" else "on-site SCC model of the harness (no gamma / Ewald), ") + ("modified Broyden mixing with the reference's defaults, vectors on the device (dftbp_b200_mixer_*)" if mixer == "device" else "modified Broyden mixing with the reference's defaults (dftbplus_b200/mixer.py)" if mixer is not None else "Anderson mixing (5 vectors, 0.4)") +
                   ", tolerance 1e-5, at most 40 iterations, Fermi smearing kT = 0.02 Ha (the synthetic tables leave no gap); "
                   "BASELINE asks for 100 steps: --md 100"}
    del H, S, D
    hp.close()
    return out


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        reference_arm(a, rank, world)
        return

    import torch
    import torch.distributed as dist
    from dftbplus_b200 import workloads, schedule
    from dftbplus_b200.api import HotPath, new_dense

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a collective mismatch must fail, not hang the box
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    w = workloads.BY_NAME[a.workload]()
    sk, geo, _ = workloads.prepare(w)
    n_spin = w["n_spin"]
    ss, nAtom, nOrb = geo["sparseSize"], geo["nAtom"], geo["nOrb"]
    pairs = int(geo["nNeighbourSK"].sum())

    stream = torch.cuda.Stream(device=dev)
    f64 = dict(dtype=torch.float64, device=dev)
    rng = np.random.default_rng(1)

    # ONE context per GPU: full pair build, this rank's share of the force chunks
    hp = HotPath(local, 0, 1)
    hp.set_stream(stream.cuda_stream)
    sk.upload(hp)
    hp.set_geometry(geo)
    hp.set_pair_partition(rank, world)
    mo = hp.mOrb

    with torch.cuda.stream(stream):
        ham, over = torch.zeros(ss, **f64), torch.zeros(ss, **f64)
        hp.buildH0S(sk.selfegy, ham, over)
        dm = over.unsqueeze(0) * torch.from_numpy(rng.normal(size=(n_spin, ss))).to(dev)
        dm = dm.contiguous()
        edm = over * torch.from_numpy(rng.normal(size=ss)).to(dev)
        shift = torch.from_numpy(scc_shift(rng, n_spin, nAtom, mo)).to(dev)
        shift_full = torch.from_numpy(0.05 * rng.normal(size=(n_spin, nAtom, mo, mo))).to(dev)
        coords_dev = torch.from_numpy(np.ascontiguousarray(geo["coords"])).to(dev)
        deriv = torch.zeros((nAtom, 3), **f64)
        flush = torch.empty(256 * 1024 * 1024 // 8, **f64)          # > 126 MB L2
    stream.synchronize()

    def step_local():
        # a new geometry every step: the coordinates are refreshed, so the device-side pair schedule
        # (pair_sort.cu) is brought up to date inside the timed region -- its records are rewritten, and
        # every 16th update it is sorted again (the CUDA-graph replay below contains the refresh form;
        # the eager steps contain the mix)
        hp.update_coords(coords_dev)
        hp.buildH0S(sk.selfegy, ham, over)                               # K1 (all pairs)
        hp.derivative_shift(dm, edm, shift=shift, deriv=deriv)           # K5 (this rank's chunks)

    def reduce_deriv():
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(deriv)                                   # assembleChunks(env, deriv)

    def step_device():
        step_local()
        reduce_deriv()

    # ---- value: device-resident, CUDA events on the launch stream -------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(a.warmup):
        step_device()
    barrier()
    sampler.mark()
    l0 = hp.kernel_launches
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * a.steps)]
    for i in range(a.steps):
        if a.flush_l2:
            with torch.cuda.stream(stream):
                flush.zero_()                                            # evict L2 between timed steps
        ev[2 * i].record(stream)
        step_device()
        ev[2 * i + 1].record(stream)
    stream.synchronize()
    barrier()
    launches = hp.kernel_launches - l0
    ms_eager = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(a.steps)) / a.steps
    ms_eager = max_over_ranks(ms_eager)
    # The same step captured once into a CUDA graph and replayed: the ~15 launches, memsets and the
    # all-reduce of a step are short, and issuing them one by one from the host leaves gaps between
    # them.  All buffers of the step are device resident with fixed addresses, so a geometry loop can
    # replay the graph with new coordinates in coords_dev.  Falls back to the eager number.
    ms_dev, graph_note = ms_eager, "eager launches"
    graph_ok = 0
    if True:
        try:
            graph = torch.cuda.CUDAGraph()
            stream.synchronize()
            with torch.cuda.graph(graph, stream=stream, capture_error_mode="thread_local"):
                step_local()            # the kernels of this rank; the all-reduce of deriv stays an eager call
            graph_ok = 1
        except Exception as e:  # noqa: BLE001
            graph_ok, graph_note = 0, "eager launches (graph capture failed: %r)" % (e,)
        if world > 1:
            # all ranks or none (the timed loops must issue the same collectives everywhere)
            if max_over_ranks(0.0 if graph_ok else 1.0) > 0.5:
                graph_ok, graph_note = 0, "eager launches (graph capture failed on some rank)"
    if graph_ok:
        for _ in range(a.warmup):
            with torch.cuda.stream(stream):
                graph.replay()
            reduce_deriv()
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * a.steps)]
        for i in range(a.steps):
            with torch.cuda.stream(stream):
                if a.flush_l2:
                    flush.zero_()
                ev[2 * i].record(stream)
                graph.replay()
            reduce_deriv()
            ev[2 * i + 1].record(stream)
        stream.synchronize()
        barrier()
        ms_dev = max_over_ranks(sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(a.steps)) / a.steps)
        graph_note = "one CUDA graph per step (captured once, replayed)" + (
            "" if world == 1 else " + an eager NCCL all-reduce of deriv")
        ms_graph = ms_dev
        if ms_eager < ms_graph:   # both are measurements of the same step; report the faster way of launching it
            ms_dev, graph_note = ms_eager, "eager launches (%.4f ms; the CUDA-graph replay of the same step took %.4f ms)" % (ms_eager, ms_graph)
        with torch.cuda.stream(stream):
            check = deriv.clone()
        step_device()
        stream.synchronize()
        if not bool(torch.allclose(check, deriv, rtol=0.0, atol=1e-9)):
            ms_dev, graph_note = ms_eager, "eager launches (graph replay did not reproduce the eager result: max diff %.3e, %d non-finite)" % (
                float((check - deriv).abs().nan_to_num(nan=1e300).max()), int((~torch.isfinite(check)).sum()))
    value = pairs / (ms_dev * 1e-3)

    # ---- per-routine device times (library timers = the reference's timer names) ----------------
    hp.timer_enable(True)
    for i in range(a.steps):
        if a.flush_l2:
            with torch.cuda.stream(stream):
                flush.zero_()
        step_device()
    t = hp.timer_read()
    # keep the GPU busy until nvidia-smi has delivered samples from inside the window.  LOCAL work
    # only: the number of iterations differs from rank to rank, so no collective may be in here.
    for i in range(1500):
        if sum(1 for ts, _ in sampler.lines if ts >= sampler.t0) >= 4 and i >= 20:
            break
        hp.buildH0S(sk.selfegy, ham, over)
        hp.derivative_shift(dm, edm, shift=shift, deriv=deriv)
    stream.synchronize()
    clocks = sampler.stop()
    k1_ms_in_step = t["sparseH0S"][0] / max(1, t["sparseH0S"][1])
    k5_ms = t["forceCalc"][0] / max(1, t["forceCalc"][1])
    # The pair build on its own: K launches back to back on the launch stream, one event pair around all
    # of them -- the GPU, not the host's launch rate or the event pair around a single short call, sets
    # the time (the library timer above brackets ONE call of two kernels with events and sees the idle
    # gap between them when the host is the slower side).  Each call writes 130 MB (> L2) afresh.
    hp.timer_enable(False)
    reps_k1 = max(20, a.steps)
    for _ in range(3):
        hp.buildH0S(sk.selfegy, ham, over)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps_k1):
        hp.buildH0S(sk.selfegy, ham, over)
    e1.record(stream)
    e1.synchronize()
    k1_ms = e0.elapsed_time(e1) / reps_k1
    hbm_peak, peak_src = peaks()
    k1_bytes = algorithmic_bytes_k1(geo)
    tr = ncu_traffic("k1_rotate_store")
    roofline = {"kernel": "pair build H0+S = k1_interp + k1_rotate_store (TMA bulk stores)", "bound": "hbm",
                "achieved": k1_bytes / k1_ms / 1e6, "peak": hbm_peak, "unit": "GB/s",
                "frac": k1_bytes / k1_ms / 1e6 / hbm_peak,
                "traffic": (tr or {}).get("bytes_per_launch"), "traffic_source": (tr or {}).get("source"),
                "traffic_note": "k1_rotate_store alone (the kernel that writes H0 and S)",
                "peak_source": peak_src, "algorithmic_bytes_per_launch": k1_bytes,
                "ms_per_launch": k1_ms, "launches_timed": reps_k1,
                "timing": "CUDA events around %d back-to-back calls on the launch stream (two kernels per call); "
                          "every call rewrites 130 MB, more than the L2 holds" % reps_k1,
                "ms_inside_step_library_timer": k1_ms_in_step}
    # what plain element-wise kernels reach on arrays of THIS size (65 MB each): torch kernels with the same
    # read : write mix as the streaming kernels of the path, L2 flushed before each -- the practical ceiling
    # next to the 2 GB copy peak of MEASURED_PEAKS.json (short kernels pay ramp-up and tail)
    stream_mix = None
    if True:
        with torch.cuda.stream(stream):
            ta, tb, tc = torch.randn(ss, **f64), torch.randn(ss, **f64), torch.randn(ss, **f64)
            to, to2 = torch.empty(ss, **f64), torch.empty((2, ss), **f64)

            def timed(fn, reps=10):
                for _ in range(2):
                    fn()
                tot = 0.0
                for _ in range(reps):
                    flush.zero_()
                    x0, x1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    x0.record(stream)
                    fn()
                    x1.record(stream)
                    x1.synchronize()
                    tot += x0.elapsed_time(x1)
                return tot / reps
            t31 = timed(lambda: torch.addcmul(ta, tb, tc, out=to))
            t11 = timed(lambda: to.copy_(ta))
            t12 = timed(lambda: torch.mul(ta.unsqueeze(0).expand(2, ss), 0.5, out=to2))
            t20 = timed(lambda: torch.dot(ta, tb))
            stream_mix = {"elements": ss,
                          "read3_write1_GBs": 32.0 * ss / t31 / 1e6, "read1_write1_GBs": 16.0 * ss / t11 / 1e6,
                          "read1_write2_GBs": 24.0 * ss / t12 / 1e6, "read2_GBs": 16.0 * ss / t20 / 1e6,
                          "note": "torch.addcmul / copy_ / broadcast mul / dot on float64 arrays of sparseSize elements"}
            del ta, tb, tc, to, to2
    kernels = {"sparseH0S": {"ms": k1_ms, "pairs_per_s": pairs / k1_ms * 1e3},
               "forceCalc": {"ms": k5_ms, "pairs_per_s_this_rank": pairs / world / k5_ms * 1e3,
                             "shift": "diagonal blocks (collinear-spin SCC)"}}
    # the same force call with full (DFTB+U-like) shift blocks: per-pair matrix products in the weights kernel
    hp.timer_enable(True)
    for i in range(3):
        hp.derivative_shift(dm, edm, shift=shift_full, deriv=deriv)
    tf = hp.timer_read()
    kernels["forceCalc_full_shift_blocks"] = {"ms": tf["forceCalc"][0] / max(1, tf["forceCalc"][1])}

    # ---- e2e: HOST (pinned) inputs and outputs, copies inside the timed region --------------------
    # N = 1: the C-ABI calls take the host pointers themselves.  N > 1: every rank uploads 1/N of
    # rho / Erho / shift and the GPUs all-gather them over NVLink, every rank downloads its 1/N
    # segment of H0 / S (the reference's ranks hold their atom chunks before assembleChunks) and
    # the all-reduced deriv.
    def pinned(shape):
        return torch.empty(shape, dtype=torch.float64, pin_memory=True)
    h_coords = pinned(geo["coords"].shape)
    h_coords.numpy()[...] = geo["coords"]
    n_in = n_spin * ss + ss + n_spin * nAtom * mo * mo
    sl_in = (n_in + world - 1) // world
    h_in = pinned(sl_in * world)
    flat_in = torch.cat([dm.reshape(-1), edm.reshape(-1), shift.reshape(-1)]).cpu()
    h_in[:n_in] = flat_in
    sl_out = (ss + world - 1) // world
    lo_out, hi_out = min(ss, rank * sl_out), min(ss, (rank + 1) * sl_out)
    h_ham, h_over, h_deriv = pinned(hi_out - lo_out), pinned(hi_out - lo_out), pinned((nAtom, 3))
    h2d = h_coords.numel() * 8 + sl_in * 8 + 8 * hp.mShell * len(sk.ang_shells)
    d2h = 2 * (hi_out - lo_out) * 8 + h_deriv.numel() * 8
    # The copies run on their own streams so that the upload of rho / Erho / shift overlaps the pair
    # schedule and the H0 / S build (which only need the coordinates) and the download of H0 / S
    # overlaps the force contraction; the C-ABI calls take device pointers on the library's stream.
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        d_in = torch.empty(sl_in * world, **f64)
        d_coords = torch.empty_like(coords_dev)
        ham_e, over_e, deriv_e = torch.empty_like(ham), torch.empty_like(over), torch.empty_like(deriv)
    g_dm = d_in[:n_spin * ss].view(n_spin, ss)
    g_edm = d_in[n_spin * ss:n_spin * ss + ss]
    g_shift = d_in[n_spin * ss + ss:n_in].view(n_spin, nAtom, mo, mo)
    stream.synchronize()

    def step_host():
        ev_c, ev_in = torch.cuda.Event(), torch.cuda.Event()
        ev_built, ev_done = torch.cuda.Event(), torch.cuda.Event()
        s_in.wait_stream(stream)                      # the previous step has consumed d_in / d_coords
        with torch.cuda.stream(s_in):
            d_coords.copy_(h_coords, non_blocking=True)
            ev_c.record(s_in)
            d_in[rank * sl_in:(rank + 1) * sl_in].copy_(h_in[rank * sl_in:(rank + 1) * sl_in], non_blocking=True)
            if world > 1:
                dist.all_gather_into_tensor(d_in, d_in[rank * sl_in:(rank + 1) * sl_in])
            ev_in.record(s_in)
        stream.wait_event(ev_c)
        stream.wait_stream(s_out)                     # the previous step's downloads have read ham_e / over_e
        hp.update_coords(d_coords)
        hp.buildH0S(sk.selfegy, ham_e, over_e)
        ev_built.record(stream)
        stream.wait_event(ev_in)
        hp.derivative_shift(g_dm, g_edm, shift=g_shift, deriv=deriv_e)
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(deriv_e)
        ev_done.record(stream)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_built)
            h_ham.copy_(ham_e[lo_out:hi_out], non_blocking=True)
            h_over.copy_(over_e[lo_out:hi_out], non_blocking=True)
            s_out.wait_event(ev_done)
            h_deriv.copy_(deriv_e, non_blocking=True)
        s_out.synchronize()                           # the step's results are on the host
    e2e_steps = max(3, min(a.steps, 10))
    for _ in range(2):
        step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_host()
    torch.cuda.synchronize()
    ms_e2e = max_over_ranks((time.perf_counter() - t0) / e2e_steps * 1e3)
    e2e = {"value": pairs / (ms_e2e * 1e-3), "unit": "pairs/s", "ms_per_step": ms_e2e,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "timing": "host wall clock, pinned host buffers, results on the host at the end of every step; uploads, kernels and "
                     "downloads on three streams (C-ABI calls with device pointers)" +
                     ("" if world == 1 else "; per rank 1/N of rho, Erho, shift up + NVLink all-gather, 1/N of H0 and S down, "
                                            "deriv all-reduced")}
    if world == 1:
        # the plain form a Fortran host would use first: synchronous C-ABI calls with HOST pointers
        np_dm = h_in[:n_spin * ss].numpy().reshape(n_spin, ss)
        np_edm = h_in[n_spin * ss:n_spin * ss + ss].numpy()
        np_shift = h_in[n_spin * ss + ss:n_in].numpy().reshape(n_spin, nAtom, mo, mo)

        def step_sync():
            hp.update_coords(h_coords.numpy())
            hp.buildH0S(sk.selfegy, h_ham.numpy(), h_over.numpy())
            hp.derivative_shift(np_dm, np_edm, shift=np_shift, deriv=h_deriv.numpy())
        for _ in range(2):
            step_sync()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_sync()
        torch.cuda.synchronize()
        ms_sync = (time.perf_counter() - t0) / e2e_steps * 1e3
        e2e["synchronous_host_pointer_calls"] = {"value": pairs / (ms_sync * 1e-3), "ms_per_step": ms_sync}

    roofline_k5 = None
    try:
        fl = algorithmic_flops_k5(geo, sk, n_spin)
        roofline_k5 = {"kernel": "k5_derivative_shift + SCC weights (this rank's share)", "bound": "fp64",
                       "flops_all_pairs": fl, "ms_per_call": k5_ms, "achieved": fl / world / k5_ms / 1e9, "unit": "TFLOP/s",
                       "note": "FP64 CUDA-core bound by design (SURVEY 8(d)); peak: the cuBLAS DGEMM rate in roofline_dm"}
    except Exception as e:  # noqa: BLE001
        roofline_k5 = {"error": repr(e)}

    # ---- N > 1: the reference's decomposition of the pair build, measured ---------------------------
    multi = None
    if world > 1 and not a.pair_only:
        hps = HotPath(local, rank, world)
        hps.set_stream(stream.cuda_stream)
        sk.upload(hps)
        hps.set_geometry(geo)
        with torch.cuda.stream(stream):
            ham2, over2 = torch.zeros(ss, **f64), torch.zeros(ss, **f64)

        def split_build():
            hps.buildH0S(sk.selfegy, ham2, over2)                 # this rank's atom chunk, rest zero
            with torch.cuda.stream(stream):
                dist.all_reduce(ham2)                             # assembleChunks (schedule.F90:170-182)
                dist.all_reduce(over2)
        for _ in range(3):
            split_build()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(10):
            split_build()
        e1.record(stream)
        stream.synchronize()
        split_ms = max_over_ranks(e0.elapsed_time(e1) / 10)
        same = bool(torch.equal(ham2, ham) and torch.equal(over2, over))
        multi = {"pair_build_replicated_ms": k1_ms, "pair_build_split_plus_allreduce_ms": split_ms,
                 "split_equals_replicated_bitwise": same,
                 "note": "split = contiguous atom chunks (getChunkRanges) + NCCL all-reduce of the zero-padded H0 and S "
                         "like assembleChunks; the faster one (replicated) is what `value` uses"}
        hps.close()
        del ham2, over2

    # ---- SCC step (dense section): (k, spin) sharded over ranks ----------------------------------
    scc, roofline_dm = None, None
    if not a.no_dense and not a.pair_only:
        # (k, spin) pairs round-robin over ranks (TParallelKS); with more GPUs than pairs every pair
        # gets a group of GPUs and its density matrix is split into column panels (schedule.dm_panels)
        n_units = n_spin
        if world <= n_units:
            my_ks, my_panel, group_size = schedule.parallel_ks(world, rank, 1, n_spin), None, 1
        else:
            group_size = world // n_units
            unit = rank // group_size
            my_ks = [(1, unit + 1)] if unit < n_units else []
            my_panel = schedule.dm_panels(geo["iAtomStart"], group_size)[rank % group_size] if my_ks else None
        n_elec = float(sum(sum(sk.occ[s - 1]) for s in w["species0"]))
        n_occ = int(round(n_elec / 2))
        with torch.cuda.stream(stream):
            # FP64 tensor-pipe denominator: cuBLAS DGEMM 8192^3 measured here (MEASURED_PEAKS.json
            # carries only bf16)
            A = torch.randn(8192, 8192, **f64)
            B = torch.randn(8192, 8192, **f64)
            for _ in range(2):
                torch.matmul(A, B)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(3):
                torch.matmul(A, B)
            e1.record(stream)
            e1.synchronize()
            dgemm_tflops = 3 * 2 * 8192.0 ** 3 / e0.elapsed_time(e1) / 1e9
            del A, B
            sq = new_dense(nOrb, device=dev)
            C = new_dense(nOrb, device=dev)
            C.copy_(torch.randn(nOrb, nOrb, **f64).mul_(nOrb ** -0.5))     # stand-in eigenvectors
            fill = torch.zeros(nOrb, **f64)
            fill[:n_occ] = 2.0 / n_spin
            # cuBLAS DSYRK on the shape of the density-matrix contraction (n x nLevels panel of C): what a
            # library call would give for the same n^2 * nLevels flops
            dsyrk_ms = None
            if my_panel is None:
                try:
                    dsyrk_ms = cublas_dsyrk_ms(stream, nOrb, n_occ, C.data_ptr(), C.stride(1) if C.stride(0) == 1 else C.stride(0),
                                               sq.data_ptr(), sq.stride(1) if sq.stride(0) == 1 else sq.stride(0))
                except Exception:  # noqa: BLE001
                    dsyrk_ms = None
            rho = torch.zeros((n_spin, ss), **f64)
            qq = torch.zeros((nAtom, mo), **f64)
            ham_scc = torch.zeros((n_spin, ss), **f64)
        stream.synchronize()
        if roofline_k5 and "achieved" in roofline_k5:
            roofline_k5["peak"] = dgemm_tflops
            roofline_k5["frac"] = roofline_k5["achieved"] / dgemm_tflops
        hp.timer_enable(True)
        reps = 2
        for rep in range(reps + 1):
            if rep == 1:
                hp.timer_enable(True)          # first pass = warm-up
                barrier()
                ts0 = torch.cuda.Event(enable_timing=True); ts1 = torch.cuda.Event(enable_timing=True)
                ts0.record(stream)
            with torch.cuda.stream(stream):
                rho.zero_()
                qq.zero_()
            # getSccHamiltonian (hamiltonian.F90:382-392): ham = 0; addShift(..., isInputZero); ham(:,1) += H0
            hp.addShift(ham_scc, over, shift, is_input_zero=True)
            with torch.cuda.stream(stream):
                ham_scc[0].add_(ham)
            for (iK, iS) in my_ks:
                hp.unpackHS(ham_scc[iS - 1], square=sq)                      # H (this spin)
                hp.unpackHS(over, square=sq)                                 # S
                # (the eigensolve runs here: H, S -> eigenvectors; timed separately below)
                if my_panel is None:
                    hp.getDensityMatrix(C, fill, dm=sq)
                    hp.packHS(rho[iS - 1], sq)
                elif my_panel[0] <= my_panel[1]:
                    hp.getDensityMatrix(C, fill, dm=sq, cols=(my_panel[2], my_panel[3]))
                    hp.packHS(rho[iS - 1], sq, atoms=(my_panel[0], my_panel[1]))
            if world > 1:
                with torch.cuda.stream(stream):
                    dist.all_reduce(rho)                                     # main.F90:4179
            for iS in range(n_spin):
                if iS % world == rank:                                       # spin channels dealt out
                    hp.mulliken(over, rho[iS], qq=qq)
            if world > 1:
                with torch.cuda.stream(stream):
                    dist.all_reduce(qq)
        ts1.record(stream)
        stream.synchronize()
        scc_ms = max_over_ranks(ts0.elapsed_time(ts1) / reps)
        t = hp.timer_read()
        per = {k: (v[0] / v[1] if v[1] else None) for k, v in t.items()}
        dm_ms = per["densityMatrix"]
        panel_frac = 1.0
        if my_panel is not None and my_panel[0] <= my_panel[1]:          # this rank's share of the triangle
            c0, c1 = my_panel[2] - 1, my_panel[3]
            panel_frac = (c1 - c0) * (2.0 * nOrb - c0 - c1) / (float(nOrb) * nOrb)
        if dm_ms:
            flops = float(nOrb) * nOrb * n_occ * panel_frac          # n^2 * nLevels, one triangle, FMA = 2 flops
            roofline_dm = {"kernel": "k3_dm_real_tma (DMMA m8n8k4; operand panels by cp.async.bulk + mbarrier pipeline, producer warp)", "bound": "tensor",
                           "achieved": flops / dm_ms / 1e9, "peak": dgemm_tflops, "unit": "TFLOP/s",
                           "frac": flops / dm_ms / 1e9 / dgemm_tflops, "traffic": None,
                           "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (FP64 is not in MEASURED_PEAKS.json)",
                           "n": nOrb, "nLevels": n_occ, "ms_per_launch": dm_ms,
                           "cublas_dsyrk_same_shape_ms": dsyrk_ms,
                           "cublas_dsyrk_same_shape_tflops": (float(nOrb) * nOrb * n_occ / dsyrk_ms / 1e9) if dsyrk_ms else None,
                           "speed_vs_cublas_dsyrk": (dsyrk_ms / dm_ms) if dsyrk_ms else None}
        gbs = lambda b, ms: (b / ms / 1e6) if ms else None
        kernels.update({
            "sparseToDense": {"ms": per["sparseToDense"], "GBs": gbs(8.0 * ss + 8.0 * nOrb * nOrb, per["sparseToDense"])},
            # packHS of a column panel only touches the blocks of the panel's atoms
            "denseToSparse": {"ms": per["denseToSparse"],
                              "GBs": gbs(24.0 * ss * (panel_frac if my_panel is not None else 1.0), per["denseToSparse"]),
                              "bytes": "24 * sparseSize" + (" * this rank's share of the blocks (approx.)" if my_panel is not None else "")},
            "densityMatrix": {"ms": dm_ms, "TFLOPs": roofline_dm["achieved"] if roofline_dm else None},
        })
        if per.get("addShift"):
            # isInputZero: over read once, ham written per spin (not read)
            kernels["addShift"] = {"ms": per["addShift"], "GBs": gbs(8.0 * ss * (1 + n_spin), per["addShift"]),
                                   "bytes": "8 * sparseSize * (1 + nSpin), isInputZero"}
        mk = t["mulliken"]
        if mk[1]:
            kernels["mulliken"] = {"ms": mk[0] / mk[1], "GBs": 16.0 * ss / (mk[0] / mk[1]) / 1e6}
        for kname in ("sparseToDense", "denseToSparse", "addShift", "mulliken"):
            if kernels.get(kname, {}).get("GBs"):
                kernels[kname]["frac_of_copy_peak"] = kernels[kname]["GBs"] / hbm_peak
        scc = {"ms_without_eigensolve": scc_ms, "ks_pairs_this_rank": len(my_ks), "gpus_per_ks_pair": group_size,
               "dm_panel_this_rank": list(my_panel) if my_panel is not None else None, "n_spin": n_spin, "nOrb": nOrb,
               "eigensolve_ms": None, "cublas_dgemm_tflops": dgemm_tflops,
               "note": "addShift (SCC Hamiltonian), then unpack(H), unpack(S), density matrix, pack per local (k,spin); all-reduce rho; Mulliken per spin"}
        del sq, C
        torch.cuda.empty_cache()
        # the eigensolve at the full size, once per (k,spin) of this rank, from a torch-free child process
        # (tools/eig_fullsize.py says why): replicas only -- every GPU of a (k,spin) group would run it
        if rank == 0 and not a.no_eigensolve:
            try:
                # torchrun exports OMP_NUM_THREADS=1; cuSOLVER's divide and conquer has a host-side part that then
                # runs three times slower (63 s instead of 20 s at n = 31744): give the child this rank's share of the cores
                env = dict(os.environ)
                env["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or world) // world))
                r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "eig_fullsize.py"), a.workload, str(local)],
                                   capture_output=True, text=True, timeout=600, env=env)
                last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
                if r.returncode == 0 and last:
                    eg = json.loads(last[-1])
                    scc["eigensolve"] = eg
                    scc["eigensolve_ms"] = eg["ms"]
                    scc["ms_with_eigensolve"] = scc_ms + eg["ms"] * len(my_ks)
                    scc["eigensolve_note"] = ("one generalized eigensolve per local (k,spin) pair, not split across the GPUs of a "
                                              "group (replicas only); measured on rank 0, added to this rank's step")
                else:
                    scc["eigensolve"] = {"error": "exit %d: %s" % (r.returncode, (r.stderr or r.stdout).strip().splitlines()[-1:])}
            except Exception as e:  # noqa: BLE001
                scc["eigensolve"] = {"error": repr(e)}
    barrier()

    # ---- BASELINE config 3 (k-points, all ranks) and config 5 (MD, rank 0 at N = 1) ---------------
    kp_sec, md_sec, md_ref_sec = None, None, None
    if not a.no_kpoints and not a.pair_only:
        try:
            kp_sec = section_kpoints(dev, stream, rank, world, dist, max_over_ranks, barrier)
        except Exception as e:  # noqa: BLE001
            kp_sec = {"error": repr(e)}
            if world > 1:
                raise
    if a.md > 0 and world == 1 and not a.pair_only:
        try:
            md_sec = section_md(dev, stream, a.md)
        except Exception as e:  # noqa: BLE001
            md_sec = {"error": repr(e)}
        try:
            md_ref_sec = section_md(dev, stream, min(a.md, 3), potential="reference")
        except Exception as e:  # noqa: BLE001
            md_ref_sec = {"error": repr(e)}

    # ---- CPU baseline (oracle port) on rank 0, N = 1 ----------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.pair_only:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import oracle, case
        oracle.set_num_threads()
        _, _, geo_o, model, selfegy = case(a.workload)
        dm_h, edm_h, sh_h = dm.cpu().numpy(), edm.cpu().numpy(), shift.cpu().numpy()

        def cpu_step():
            oracle.build_h0(model, geo_o, selfegy)
            oracle.build_s(model, geo_o)
            oracle.derivative_shift(model, geo_o, dm_h, edm_h, shift=sh_h)
        cpu_step()
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            cpu_step()
            best = min(best, time.perf_counter() - t0)
        cpu = {"value": pairs / best, "unit": "pairs/s", "cores": oracle.num_threads(), "kind": "port",
               "ms_per_step": best * 1e3,
               "sample": "the full workload once (best of 3): buildH0 + buildS + derivative_shift(SCC, %d spins); "
                         "C/OpenMP restatement of the reference CPU path, gcc -O2 -funroll-all-loops -fopenmp" % n_spin}

    # ---- neighbour list on the device, N = 1 only, in a child process -------------
    n2 = None
    if rank == 0 and world == 1 and not a.no_n2 and not a.pair_only:
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "n2_probe.py"), a.workload, "--oracle"],
                               capture_output=True, text=True, timeout=300)
            last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            n2 = json.loads(last[-1]) if r.returncode == 0 and last else {
                "error": "exit %d: %s" % (r.returncode, (r.stderr or r.stdout).strip().splitlines()[-1:])}
        except Exception as e:  # noqa: BLE001  (timeout included)
            n2 = {"error": repr(e)}

    # ---- Coulomb 1/R matrix on the device (second piece), N = 1 only, in a child process ----
    coul = None
    if rank == 0 and world == 1 and not a.no_n2 and not a.pair_only:
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "coulomb_probe.py"), a.workload, "--oracle"],
                               capture_output=True, text=True, timeout=300)
            last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            coul = json.loads(last[-1]) if r.returncode == 0 and last else {
                "error": "exit %d: %s" % (r.returncode, (r.stderr or r.stdout).strip().splitlines()[-1:])}
        except Exception as e:  # noqa: BLE001  (timeout included)
            coul = {"error": repr(e)}

    if rank == 0:
        cfg = config_for(a.workload, geo, n_spin, a.flush_l2)
        line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_dev, "ms_per_step_eager": ms_eager,
                "launch_mode": graph_note, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg,
                "parallelism": "one context per GPU: pair build replicated, force chunks dealt out round-robin + all-reduce of deriv, "
                               "(k,spin) round-robin, density-matrix column panels beyond",
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "roofline_k5": roofline_k5, "roofline_dm": roofline_dm, "kernels": kernels, "stream_mix_same_size": stream_mix,
                "scc_step": scc,
                "kpoints_si512": kp_sec, "md_water2046": md_sec, "md_water2046_reference_potential": md_ref_sec, "multi_gpu": multi, "cpu_baseline": cpu,
                "neighbour_list": n2, "coulomb_ewald": coul}
        print(json.dumps(line), flush=True)
    hp.close()
    if world > 1:
        # all ranks are done with their collectives here; leave without tearing NCCL down (a rank that is
        # still inside destroy_process_group when its peers are gone can hang the launcher)
        barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == "__main__":
    main()
