#!/usr/bin/env python3
"""bench.py -- SK pairs/s for the H0/S build + force contraction (BASELINE.json metric) and the
SCC-step kernel times, on N B200s of one node.

A step = one pass of the pair hot path over one synthetic geometry: buildH0 + buildS (one fused
launch) and derivative_shift (SCC form, collinear spin) for every neighbour pair of the workload.
  value  device-resident: inputs in HBM, CUDA events on the launch stream, max over ranks;
  e2e    the same step through the C ABI with HOST (pinned) buffers: every array crosses PCIe.
Extra objects: roofline (K1, HBM bound), roofline_dm (K3, FP64 tensor pipe), kernels (device time
per routine with the reference's timer names), scc_step (unpack -> [cuSOLVER] -> DM -> pack ->
Mulliken per spin), cpu_baseline (the oracle = restatement of the reference CPU path, rank 0, N=1).

`--impl reference` times that CPU restatement itself (the reference is Fortran and cannot be
built in this image: no Fortran compiler) with all host threads.

Multi-GPU (one process per GPU, torchrun): the pair build is replicated (it is cheaper than moving
130 MB of sparse matrix over NVLink -- "replicas only" for K1), the force contraction is split over
ranks by the reference's workload-balanced atom sets with an NCCL all-reduce of deriv(3,nAtom), and
the (k-point, spin) dense work is sharded round-robin with an NCCL all-reduce of the sparse rho.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="fec4096")
    ap.add_argument("--no-dense", action="store_true", help="skip the SCC-step (dense) section")
    ap.add_argument("--eigensolve", action="store_true", help="include cuSOLVER at the full workload size")
    ap.add_argument("--no-n2", action="store_true", help="skip the neighbour-list report (tools/n2_probe.py)")
    return ap.parse_args()


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
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*ncu_all_kernels*.txt")))
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


def scc_shift(rng, n_spin, n_atom, m_orb):
    """Block shift of a collinear-spin SCC run without DFTB+U / on-site corrections (BASELINE config
    4): diagonal mOrb x mOrb blocks, one value per orbital (hamiltonian.F90:148-293 -> shift.F90:232-309
    expand atom / shell potentials onto the block diagonal)."""
    sh = np.zeros((n_spin, n_atom, m_orb, m_orb))
    idx = np.arange(m_orb)
    sh[:, :, idx, idx] = 0.05 * rng.normal(size=(n_spin, n_atom, m_orb))
    return sh


def reference_arm(a, rank, world):
    """The reference's CPU path (oracle restatement, C + OpenMP, all host threads): same step."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import oracle, case
    oracle.set_num_threads()          # all host cores (torchrun exports OMP_NUM_THREADS=1)
    w, sk, geo, model, selfegy = case(a.workload)
    rng = np.random.default_rng(1)
    s = oracle.build_s(model, geo)
    dm = s * rng.normal(size=(2, geo["sparseSize"]))
    edm = s * rng.normal(size=geo["sparseSize"])
    shift = scc_shift(rng, 2, geo["nAtom"], model.mOrb)
    pairs = int(geo["nNeighbourSK"].sum())

    def step():
        oracle.build_h0(model, geo, selfegy)
        oracle.build_s(model, geo)
        oracle.derivative_shift(model, geo, dm, edm, shift=shift)
    for _ in range(min(a.warmup, 1)):
        step()
    steps = max(1, min(a.steps, 5))          # bounded: about 1.5 s of 16-thread CPU work per step
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    val = pairs / dt
    line = {"impl": "reference", "metric": "sk_pairs_per_s_h0s_plus_force", "value": val, "unit": "pairs/s",
            "n_gpus": a.gpus, "steps": steps, "warmup": min(a.warmup, 1), "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": a.workload, "nAtom": int(geo["nAtom"]), "nOrb": int(geo["nOrb"]), "pairs": pairs,
                       "sparseSize": int(geo["sparseSize"]), "n_spin": 2, "sk_tables": "synthetic .skf, 0.02 bohr x 500"},
            "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": oracle.num_threads(), "kind": "port",
                             "sample": "full workload, %d steps: buildH0 + buildS + derivative_shift(SCC, 2 spins)" % steps},
            "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "restatement of the reference CPU path (OpenMP over atoms), not a DFTB+ build"}
    print(json.dumps(line))


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

    # full-build context (K1 replicated on every rank) and a sharded one for forces / Mulliken
    hp = HotPath(local, 0, 1)
    hp.set_stream(stream.cuda_stream)
    sk.upload(hp)
    hp.set_geometry(geo)
    hpw = hp
    if world > 1:
        hpw = HotPath(local, rank, world)
        hpw.set_stream(stream.cuda_stream)
        sk.upload(hpw)
        hpw.set_geometry(geo)

    with torch.cuda.stream(stream):
        ham, over = torch.zeros(ss, **f64), torch.zeros(ss, **f64)
        hp.buildH0S(sk.selfegy, ham, over)
        dm = over.unsqueeze(0) * torch.from_numpy(rng.normal(size=(n_spin, ss))).to(dev)
        dm = dm.contiguous()
        edm = over * torch.from_numpy(rng.normal(size=ss)).to(dev)
        shift = torch.from_numpy(scc_shift(rng, n_spin, nAtom, hp.mOrb)).to(dev)
        shift_full = torch.from_numpy(0.05 * rng.normal(size=(n_spin, nAtom, hp.mOrb, hp.mOrb))).to(dev)
        coords_dev = torch.from_numpy(np.ascontiguousarray(geo["coords"])).to(dev)
        deriv = torch.zeros((nAtom, 3), **f64)
        flush = torch.empty(256 * 1024 * 1024 // 8, **f64)          # > 126 MB L2
    stream.synchronize()

    def step_device():
        # a new geometry every step: the coordinates are refreshed, so the device-side pair
        # schedule (pair_sort.cu: histogram, scan, scatter) is rebuilt inside the timed region
        hp.update_coords(coords_dev)
        if hpw is not hp:
            hpw.update_coords(coords_dev)
        hp.buildH0S(sk.selfegy, ham, over)                               # K1 (all pairs)
        hpw.derivative_shift(dm, edm, shift=shift, deriv=deriv)          # K5 (this rank's atoms)
        if world > 1:
            with torch.cuda.stream(stream):
                dist.all_reduce(deriv)                                   # assembleChunks(env, deriv)

    # ---- value: device-resident, CUDA events on the launch stream -------------------------------
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(a.warmup):
        step_device()
    barrier()
    sampler.mark()
    l0 = hp.kernel_launches + (hpw.kernel_launches if hpw is not hp else 0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * a.steps)]
    for i in range(a.steps):
        with torch.cuda.stream(stream):
            flush.zero_()                                                # evict L2 between timed steps
        ev[2 * i].record(stream)
        step_device()
        ev[2 * i + 1].record(stream)
    stream.synchronize()
    barrier()
    launches = hp.kernel_launches + (hpw.kernel_launches if hpw is not hp else 0) - l0
    ms_dev = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(a.steps)) / a.steps
    ms_dev = max_over_ranks(ms_dev)
    value = pairs / (ms_dev * 1e-3)

    # ---- per-routine device times (library timers = the reference's timer names) ----------------
    hp.timer_enable(True)
    if hpw is not hp:
        hpw.timer_enable(True)
    for i in range(a.steps):
        with torch.cuda.stream(stream):
            flush.zero_()
        step_device()
    t = hp.timer_read()
    tw = hpw.timer_read()
    # keep the GPU busy until nvidia-smi has delivered samples from inside the window.  LOCAL work
    # only: the number of iterations differs from rank to rank, so no collective may be in here.
    for i in range(1500):
        if sum(1 for ts, _ in sampler.lines if ts >= sampler.t0) >= 4 and i >= 20:
            break
        hp.buildH0S(sk.selfegy, ham, over)
        hpw.derivative_shift(dm, edm, shift=shift, deriv=deriv)
    stream.synchronize()
    clocks = sampler.stop()
    k1_ms = t["sparseH0S"][0] / max(1, t["sparseH0S"][1])
    k5_ms = tw["forceCalc"][0] / max(1, tw["forceCalc"][1])
    hbm_peak, peak_src = peaks()
    k1_bytes = algorithmic_bytes_k1(geo)
    roofline = {"kernel": "k1_pair_build<H0+S>", "bound": "hbm", "achieved": k1_bytes / k1_ms / 1e6,
                "peak": hbm_peak, "unit": "GB/s", "frac": k1_bytes / k1_ms / 1e6 / hbm_peak,
                "traffic": (ncu_traffic("k1_pair_build") or {}).get("bytes_per_launch"),
                "traffic_source": (ncu_traffic("k1_pair_build") or {}).get("source"),
                "peak_source": peak_src, "algorithmic_bytes_per_launch": k1_bytes,
                "ms_per_launch": k1_ms, "l2": "flushed between steps (256 MB memset)"}
    kernels = {"sparseH0S": {"ms": k1_ms, "pairs_per_s": pairs / k1_ms * 1e3},
               "forceCalc": {"ms": k5_ms, "pairs_per_s_this_rank": hpw.num_pairs / k5_ms * 1e3,
                             "shift": "diagonal blocks (collinear-spin SCC)"}}
    # the same force call with full (DFTB+U-like) shift blocks: per-pair matrix products in the weights kernel
    hpw.timer_enable(True)
    for i in range(3):
        hpw.derivative_shift(dm, edm, shift=shift_full, deriv=deriv)
    tf = hpw.timer_read()
    kernels["forceCalc_full_shift_blocks"] = {"ms": tf["forceCalc"][0] / max(1, tf["forceCalc"][1])}
    hp.timer_enable(True)
    if hpw is not hp:
        hpw.timer_enable(True)

    # ---- e2e: the C-ABI call with HOST buffers (pinned), copies inside the timed region ----------
    def pinned(shape):
        return torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
    h_ham, h_over = pinned(ss), pinned(ss)
    h_dm, h_edm = pinned((n_spin, ss)), pinned(ss)
    h_shift, h_deriv = pinned((n_spin, nAtom, hp.mOrb, hp.mOrb)), pinned((nAtom, 3))
    h_coords = pinned(geo["coords"].shape)
    h_dm[...] = dm.cpu().numpy(); h_edm[...] = edm.cpu().numpy(); h_shift[...] = shift.cpu().numpy()
    h_coords[...] = geo["coords"]
    h2d = h_coords.nbytes + h_dm.nbytes + h_edm.nbytes + h_shift.nbytes + 8 * hp.mShell * len(sk.ang_shells)
    d2h = h_ham.nbytes + h_over.nbytes + h_deriv.nbytes

    def step_host():
        hp.update_coords(h_coords)
        hp.buildH0S(sk.selfegy, h_ham, h_over)
        if hpw is not hp:
            hpw.update_coords(h_coords)
        hpw.derivative_shift(h_dm, h_edm, shift=h_shift, deriv=h_deriv)
        if world > 1:
            td = torch.from_numpy(h_deriv).to(dev)
            dist.all_reduce(td)
            h_deriv[...] = td.cpu().numpy()
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
           "timing": "host wall clock around synchronous C-ABI calls, pinned host buffers"}

    # ---- SCC step (dense section): (k, spin) sharded over ranks ----------------------------------
    scc, roofline_dm = None, None
    if not a.no_dense:
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
            rho = torch.zeros((n_spin, ss), **f64)
            qq = torch.zeros((nAtom, hp.mOrb), **f64)
            ham_scc = torch.zeros((n_spin, ss), **f64)
        stream.synchronize()
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
            # getSccHamiltonian (hamiltonian.F90:382-392): ham = 0; addShift(..., isInputZero); ham(:,1) += H0
            hp.addShift(ham_scc, over, shift, is_input_zero=True)
            with torch.cuda.stream(stream):
                ham_scc[0].add_(ham)
            for (iK, iS) in my_ks:
                hp.unpackHS(ham_scc[iS - 1], square=sq)                      # H (this spin)
                hp.unpackHS(over, square=sq)                                 # S
                # (cuSOLVER hegvd would run here: H, S -> eigenvectors; timed separately below)
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
                hpw.mulliken(over, rho[iS], qq=qq)
            if world > 1:
                with torch.cuda.stream(stream):
                    dist.all_reduce(qq)
        ts1.record(stream)
        stream.synchronize()
        scc_ms = max_over_ranks(ts0.elapsed_time(ts1) / reps)
        t = hp.timer_read()
        per = {k: (v[0] / v[1] if v[1] else None) for k, v in t.items()}
        dm_ms = per["densityMatrix"]
        if dm_ms:
            flops = float(nOrb) * nOrb * n_occ          # n^2 * nLevels, one triangle, FMA = 2 flops
            if my_panel is not None:                    # this rank's share of the triangle
                c0, c1 = my_panel[2] - 1, my_panel[3]
                flops *= (c1 - c0) * (2.0 * nOrb - c0 - c1) / (float(nOrb) * nOrb)
            roofline_dm = {"kernel": "k3_dm_real (DMMA m8n8k4)", "bound": "tensor",
                           "achieved": flops / dm_ms / 1e9, "peak": dgemm_tflops, "unit": "TFLOP/s",
                           "frac": flops / dm_ms / 1e9 / dgemm_tflops, "traffic": None,
                           "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (FP64 is not in MEASURED_PEAKS.json)",
                           "n": nOrb, "nLevels": n_occ, "ms_per_launch": dm_ms}
        gbs = lambda b, ms: (b / ms / 1e6) if ms else None
        kernels.update({
            "sparseToDense": {"ms": per["sparseToDense"], "GBs": gbs(8.0 * ss + 8.0 * nOrb * nOrb, per["sparseToDense"])},
            "denseToSparse": {"ms": per["denseToSparse"], "GBs": gbs(24.0 * ss, per["denseToSparse"])},
            "densityMatrix": {"ms": dm_ms, "TFLOPs": roofline_dm["achieved"] if roofline_dm else None},
        })
        if per.get("addShift"):
            kernels["addShift"] = {"ms": per["addShift"], "GBs": gbs(8.0 * ss * (1 + 2 * n_spin), per["addShift"])}
        tw = hpw.timer_read() if hpw is not hp else t
        mk = tw["mulliken"]
        if mk[1]:
            kernels["mulliken"] = {"ms": mk[0] / mk[1], "GBs": 16.0 * ss / (mk[0] / mk[1]) / 1e6}
        scc = {"ms_without_eigensolve": scc_ms, "ks_pairs_this_rank": len(my_ks), "gpus_per_ks_pair": group_size,
               "dm_panel_this_rank": list(my_panel) if my_panel is not None else None, "n_spin": n_spin, "nOrb": nOrb,
               "eigensolve_ms": None, "cublas_dgemm_tflops": dgemm_tflops,
               "note": "addShift (SCC Hamiltonian), then unpack(H), unpack(S), density matrix, pack per local (k,spin); all-reduce rho; Mulliken per spin"}
        del sq, C
        # eigensolve: cuSOLVER dsygvd, at the full size only on request (it dominates by far)
        if rank == 0:
            name = a.workload if a.eigensolve else "water1002"
            w2 = w if a.eigensolve else workloads.BY_NAME[name]()
            sk2, geo2, _ = (sk, geo, None) if a.eigensolve else workloads.prepare(w2)
            hp2 = HotPath(local)
            hp2.set_stream(stream.cuda_stream)
            sk2.upload(hp2)
            hp2.set_geometry(geo2)
            n2 = geo2["nOrb"]
            with torch.cuda.stream(stream):
                h2, s2 = torch.zeros(geo2["sparseSize"], **f64), torch.zeros(geo2["sparseSize"], **f64)
                H2, S2 = new_dense(n2, device=dev), new_dense(n2, device=dev)
            hp2.buildH0S(sk2.selfegy, h2, s2)
            hp2.timer_enable(True)
            for _ in range(2):
                hp2.unpackHS(h2, square=H2)
                hp2.unpackHS(s2, square=S2)
                hp2.diagDenseMtx(H2, S2)
            tt = hp2.timer_read()
            scc["eigensolve"] = {"workload": name, "nOrb": n2, "ms": tt["diagonalization"][0] / 2,
                                 "solver": "cusolverDnDsygvd"}
            if a.eigensolve:
                scc["eigensolve_ms"] = scc["eigensolve"]["ms"]
            hp2.close()

    # ---- CPU baseline (oracle port) on rank 0, N = 1 ----------------------------------------------
    cpu = None
    if rank == 0 and world == 1:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import oracle, case
        oracle.set_num_threads()
        _, _, geo_o, model, selfegy = case(a.workload)
        dm_h, edm_h, sh_h = h_dm.copy(), h_edm.copy(), h_shift.copy()

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
    # Not part of `value`: a separate report.  Own process = own CUDA context and a timeout, so this
    # young path can neither hang nor disturb the numbers above (all of which are final here).
    n2 = None
    if rank == 0 and world == 1 and not a.no_n2:
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "n2_probe.py"), a.workload, "--oracle"],
                               capture_output=True, text=True, timeout=300)
            last = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
            n2 = json.loads(last[-1]) if r.returncode == 0 and last else {
                "error": "exit %d: %s" % (r.returncode, (r.stderr or r.stdout).strip().splitlines()[-1:])}
        except Exception as e:  # noqa: BLE001  (timeout included)
            n2 = {"error": repr(e)}

    if rank == 0:
        line = {"metric": "sk_pairs_per_s_h0s_plus_force", "value": value, "unit": "pairs/s", "n_gpus": world,
                "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_dev, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": a.workload, "nAtom": nAtom, "nOrb": nOrb, "pairs": pairs,
                           "sparseSize": ss, "n_spin": n_spin, "sk_tables": "synthetic .skf, 0.02 bohr x 500",
                           "l2": "flushed between timed steps (256 MB memset); H0+S output 2x%.0f MB" % (ss * 8 / 1e6),
                           "parallelism": "K1 replicated, K5 atom-sharded + all-reduce, (k,spin) round-robin, "
                                          "density-matrix column panels beyond"},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
                "roofline_dm": roofline_dm, "kernels": kernels, "scc_step": scc, "cpu_baseline": cpu,
                "neighbour_list": n2}
        print(json.dumps(line))
    hp.close()
    if hpw is not hp:
        hpw.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
