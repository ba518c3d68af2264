#!/usr/bin/env python
"""Benchmark of the covariance hot path (BASELINE.json metric: K_ff/K_ef build GFLOP/s, K-build wall time,
E/F predictions per second) on 1..8 B200.

    python bench.py --gpus N --steps K --warmup W              # this framework (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU implementation (rank 0 only)

Workload (config.workload, SURVEY.md §8d S5 = BASELINE.json configs[4], the scale-out configuration, which
fits one GPU): 340 synthetic Cu fcc 2x2x2 structures (32 atoms), SO3(nmax=3, lmax=4, rcut=5.0) computed on
the device, every atom a force centre: N = 340 energy rows + 32 640 force rows, RBF zeta = 2,
(sigma, l) = (1.0, 0.1), noise 0.002 / 0.1.

A step = one likelihood-loop covariance build: K and dK/dl of the training set (K_ee, K_ef/K_fe, K_ff
with the gradient epilogue, gaussianprocess.py:158-159 -> RBF_mb.k_total_with_grad), rows block-sharded
over the ranks, K all-gathered over NCCL and mirrored (strong scaling: the total work is fixed).
  value  : algorithmic GFLOP/s = (32 d P_ff + 8 d P_ef + 2 d P_ee) / step time, P = same-species atom pairs
           actually evaluated (the symmetric build evaluates the J >= I blocks only), operands resident in HBM.
  e2e    : the same flops over the time of one GP.log_marginal_likelihood(theta, eval_gradient=True)
           call starting from HOST (pinned) packed arrays: H2D + row packing + K/dK build + cuSOLVER
           potrf/potrs + inverse rows (trailing-block potrs) + gradient trace + D2H of the reduced scalars.
  roofline: K_ff kernel alone (CUDA events around gprb_kff on its stream) against the FP64 tensor (DMMA)
           peak measured live by gprb_fp64_dmma_peak (MEASURED_PEAKS.json carries no fp64 entry).
  cpu_baseline: the UNMODIFIED reference C++ (oracle/_ref) on a bounded sample of the same workload,
           row-split over the host cores like the reference's MPI split.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

D = 30
SIGMA, ELL, ZETA = 1.0, 0.1, 2.0
NOISE_E, NOISE_F = 0.002, 0.1
WORKLOADS = {
    # name: (structures, fcc repeats, seed0, description)
    "s5": (340, 2, 2000, "S5: 340 x Cu fcc 32 atoms, all atoms force centres, N = 340 E + 32640 F rows"),
    "s4": (200, 3, 1000, "S4: 200 x Cu fcc 108 atoms, all atoms force centres, N = 200 E + 64800 F rows"),
    "medium": (100, 2, 2000, "100 x Cu fcc 32 atoms (profiling size), N = 100 E + 9600 F rows"),
    "small": (24, 2, 2000, "24 x Cu fcc 32 atoms (development size)"),
}


# DRAM bytes per launch of the K_ff(+grad) kernel, from ncu (13.82 GB read + 17.18 GB written at S5)
KFF_DRAM_TRAFFIC = {"s5": 30.995e9}


SO3_CPU_SECONDS = []     # numpy port of SO3.calculate (oracle/so3.py), seconds per structure of the CPU sample


def flops_of(p_ff, p_ef, p_ee, d=D):
    return 32.0 * d * p_ff + 8.0 * d * p_ef + 2.0 * d * p_ee


# ------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def summary(self, t0, t1):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, line in self.rows:
            if t < t0 or t > t1:
                continue
            f = [v.strip() for v in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: unmodified reference C++ (oracle/_ref), row-split over host threads
# ------------------------------------------------------------------------------------------------
def cpu_sample(n_struct, nrep, seed0):
    """Host packed training data of the first `n_struct` structures of the workload (device SO3 when a
    GPU is present is NOT used here: the sample is produced by the numpy oracle descriptor so that the
    reference arm runs without touching the product)."""
    from oracle import so3 as oso3
    from gpr_calculator_b200.synthetic import cu_fcc
    e_items, f_items = [], []
    global SO3_CPU_SECONDS
    for k in range(n_struct):
        at, _, _ = cu_fcc(nrep, seed0 + k)
        t0 = time.perf_counter()
        x, dxdr, seq = oso3.so3_calculate(at.positions, at.cell, at.pbc, at.numbers, 3, 4, 5.0, 2.0)
        SO3_CPU_SECONDS.append(time.perf_counter() - t0)
        ele = np.asarray(at.numbers, dtype=np.int32)
        e_items.append((x, ele))
        for i in range(len(at)):
            ids = np.flatnonzero(seq[:, 1] == i)
            f_items.append((x[seq[ids, 0]], dxdr[ids], ele[seq[ids, 0]]))
    return e_items, f_items


def reference_build(e_items, f_items, threads):
    """K and dK/dl blocks with grad, the way RBF_mb.k_total_with_grad assembles them (RBF_mb.py:173-204),
    each block row-split over `threads` workers like the MPI split (RBF_mb.py:257-301, 348-431, 471-481)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import kernels as ok
    O = ok.RBFOracle("ref" if ok.have_ref() else "port")
    E, F = ok.list_to_tuple(e_items, mode="energy"), ok.list_to_tuple(f_items)

    def chunks(items):
        n = len(items)
        per = max(1, -(-n // threads))
        return [items[i:i + per] for i in range(0, n, per)]

    with ThreadPoolExecutor(max_workers=threads) as pool:
        ff = list(pool.map(lambda c: O.kff_C(ok.list_to_tuple(c), F, SIGMA, ELL, ZETA, grad=True), chunks(f_items)))
        fe = list(pool.map(lambda c: O.kef_C(E, ok.list_to_tuple(c), SIGMA, ELL, ZETA, grad=True), chunks(f_items)))
        ee = list(pool.map(lambda c: O.kee_C(ok.list_to_tuple(c, mode="energy"), E, SIGMA, ELL, ZETA, grad=True), chunks(e_items)))
    Kff = np.vstack([b[0] for b in ff])
    Kef = np.hstack([b[0] for b in fe])
    Kee = np.vstack([b[0] for b in ee])
    return np.block([[Kee, Kef], [Kef.T, Kff]])


def sample_pairs(e_items, f_items):
    re_ = sum(len(x) for x, _ in e_items)
    rf = sum(len(x) for x, _, _ in f_items)
    return rf * rf, re_ * rf, re_ * re_    # single species: every row pair counts; full rectangular blocks


def run_cpu(n_struct, nrep, seed0, steps, warmup, threads):
    from oracle import kernels as ok
    ok.build(ref=os.path.isdir("/root/reference"), port=True)
    e_items, f_items = cpu_sample(n_struct, nrep, seed0)
    fl = flops_of(*sample_pairs(e_items, f_items))
    for _ in range(warmup):
        reference_build(e_items, f_items, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        reference_build(e_items, f_items, threads)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    kind = "reference" if ok.have_ref() else "port"
    sample = ("first %d structures of the workload (%d E + %d F centres, %.3g K_ff pairs), K and dK/dl "
              "(rbf_k*_many_with_grad)" % (n_struct, len(e_items), len(f_items), sample_pairs(e_items, f_items)[0]))
    return fl / dt * 1e-9, dt, kind, sample


# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="s5", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-structures", type=int, default=None, help="structures in the CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-predict", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile-e2e", default=None, help="write a cProfile listing of one end-to-end step to this file")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_struct, nrep, seed0, desc = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    config = {"workload": desc, "kernel": "RBF zeta=2 sigma=%g l=%g" % (SIGMA, ELL), "descriptor": "SO3 nmax=3 lmax=4 rcut=5.0 (d=30)",
              "step": "K and dK/dl of the training set (K_ee + K_ef + K_ff, gradient epilogue)",
              "l2": "inputs (packed rows, >300 MB) and outputs (>8 GB) exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return
        n_cpu = args.cpu_structures or 5
        gf, dt, kind, sample = run_cpu(n_cpu, nrep, seed0, args.steps, max(args.warmup, 0), threads)
        print(json.dumps({"impl": "reference", "metric": "covariance_build_gflops", "value": gf, "unit": "GFLOP/s",
                          "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3,
                          "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
                          "data": "synthetic", "config": config,
                          "cpu_baseline": {"value": gf, "unit": "GFLOP/s", "cores": threads, "kind": kind, "sample": sample},
                          "e2e": {"value": gf, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from gpr_calculator_b200 import _lib, device as gdev, dist as gdist, synthetic as syn
    from gpr_calculator_b200.SO3 import SO3
    from gpr_calculator_b200.gaussianprocess import GP
    from gpr_calculator_b200.kernels import RBF_mb
    lib = _lib.load()

    # ---- synthetic training set, descriptors on the device ------------------------------------------
    labelled = syn.structures(n_struct, nrep, seed0)
    des = SO3(nmax=3, lmax=4, rcut=5.0)
    E_dev, F_dev = syn.packed_from_batch(des, [a for a, _, _ in labelled])
    y = syn.targets(labelled)
    e_pack = gdev.Pack(E_dev[0], E_dev[1], E_dev[2])
    f_pack = gdev.Pack(F_dev[0], F_dev[2], F_dev[3], dxdr=F_dev[1])
    NE, NF = e_pack.n_groups, f_pack.n_groups
    N = NE + 3 * NF
    ele_f, ele_e = F_dev[2].cpu().numpy(), E_dev[1].cpu().numpy()
    p_ff = syn.pair_counts(ele_f, F_dev[3], symmetric=True)          # J >= I blocks only
    p_ee = syn.pair_counts(ele_e, E_dev[2], symmetric=False)
    p_ef = e_pack.pair_count(f_pack)                                  # one pass writes K_ef and K_fe
    flops = flops_of(p_ff, p_ef, p_ee)
    config.update({"N": N, "force_rows": int(F_dev[0].shape[0]), "energy_rows": int(E_dev[0].shape[0]),
                   "pairs_ff_evaluated": p_ff, "parallelism": "row-block x%d" % world,
                   "gather": "none (1 GPU)" if world == 1 else
                   ("fused into the K_fe/K_ff epilogue (NVLink peer stores, gprb_k*_multi)" if gdist.peer_gather_enabled()
                    else "NCCL all-gather (GPRB_NO_PEER=1)")})

    gp = GP(kernel=RBF_mb(para=[SIGMA, ELL], zeta=ZETA), descriptor=des, noise_e=NOISE_E, noise_f=NOISE_F, log_file=None)
    gp.train_x = {"energy": e_pack, "force": f_pack}      # device-resident packs (gdev.packs_of passes Packs through)
    gp.y_train = y

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        return gp._build_K(grad=True)

    for _ in range(max(args.warmup, 0)):
        out = step()
        del out
    clocks = Clocks(local_rank) if rank == 0 else None
    barrier()
    launches0 = lib.gprb_launch_count()
    _lib.PROFILE = []
    t_wall0 = time.time()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        out = step()
        del out
    ev1.record()
    barrier()
    t_wall1 = time.time()
    prof, _lib.PROFILE = _lib.PROFILE, None
    launches = lib.gprb_launch_count() - launches0
    ms_step = ev0.elapsed_time(ev1) / args.steps
    kff_ms = [a.elapsed_time(b) for n, a, b, _ in prof if n in ("gprb_kff", "gprb_kff_multi")]
    kff_ms_avg = float(np.mean(kff_ms))
    t = torch.tensor([ms_step, kff_ms_avg], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step, kff_ms_max = float(t[0]), float(t[1])
    clk = clocks.summary(t_wall0, t_wall1) if clocks else None

    # ---- roofline of the dominant kernel (K_ff DMMA kernel), per launch on this rank ------------------
    peak = np.zeros(1)
    _lib.call("gprb_fp64_dmma_peak", peak.ctypes.data, gdev.stream())
    if world == 1:
        my_pff = p_ff
    else:
        windows = gdist.row_windows(e_pack.indices, f_pack.indices, world, upper=True)
        (f0, f1) = windows[rank][1]
        rows = np.asarray(f_pack.indices, dtype=np.float64)
        suffix = np.cumsum(rows[::-1])[::-1]
        my_pff = float((rows[f0:f1] * suffix[f0:f1]).sum())
    achieved = 32.0 * D * my_pff / (kff_ms_avg * 1e-3) * 1e-12
    roofline = {"bound": "tensor", "kernel": "cov_mma_kernel<4,8,RBF,grad> (gprb_kff, FP64 DMMA.8x8x4)",
                "achieved": achieved, "peak": float(peak[0]), "unit": "TFLOP/s", "frac": achieved / float(peak[0]),
                "traffic": KFF_DRAM_TRAFFIC.get(args.workload) if world == 1 else None,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel at this "
                                  "size (profiles/r01_kff_s5_ncu_summary.txt); K and dK/dl written once = 17.05 GB",
                "peak_source": "live DMMA.8x8x4 issue-rate microbenchmark (gprb_fp64_dmma_peak); "
                "MEASURED_PEAKS.json has no fp64 entry; cuBLAS DGEMM 8192^3 on this pool: 35.5 TFLOP/s",
                "kff_ms_per_launch": kff_ms_avg, "algorithmic_flops_per_pair": 32 * D}

    result = {"metric": "covariance_build_gflops", "value": flops / (ms_step * 1e-3) * 1e-9, "unit": "GFLOP/s",
              "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
              "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
              "config": config, "roofline": roofline, "gpu_launches": int(launches), "clocks": clk,
              "k_build_ms": ms_step, "kff_ms_max_over_ranks": kff_ms_max}

    # ---- end to end: GP.log_marginal_likelihood from pinned host arrays ------------------------------
    if not args.no_e2e:
        E_host, keep1 = syn.to_host(E_dev, pin=True)
        F_host, keep2 = syn.to_host(F_dev, pin=True)
        h2d = sum(a.nbytes for a in E_host[:2]) + sum(a.nbytes for a in F_host[:3])
        gp2 = GP(kernel=RBF_mb(para=[SIGMA, ELL], zeta=ZETA), descriptor=des, noise_e=NOISE_E, noise_f=NOISE_F, log_file=None)
        gp2.train_x = {"energy": E_host, "force": F_host}
        gp2.y_train = y
        theta = np.array([SIGMA, ELL])

        def e2e_step():
            gdev.clear_cache()        # every step re-sends the host arrays and re-packs the rows
            return gp2.log_marginal_likelihood(theta, eval_gradient=True)

        res = None
        for _ in range(min(max(args.warmup, 0), 3)):
            res = e2e_step()
        if args.profile_e2e and rank == 0:
            import cProfile
            import io
            import pstats
            pr = cProfile.Profile()
            pr.enable()
            e2e_step()
            torch.cuda.synchronize()
            pr.disable()
            buf = io.StringIO()
            pstats.Stats(pr, stream=buf).sort_stats("cumulative").print_stats(40)
            with open(args.profile_e2e, "w") as fh:
                fh.write(buf.getvalue())
        elif args.profile_e2e:
            e2e_step()
        barrier()
        _lib.PROFILE = []
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / args.steps
        prof, _lib.PROFILE = _lib.PROFILE, None
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
        parts, host_parts = {}, {}
        for n, a, b, h in prof:
            parts[n] = parts.get(n, 0.0) + a.elapsed_time(b) / args.steps
            host_parts[n] = host_parts.get(n, 0.0) + h * 1e3 / args.steps
        # device -> host reads of a step: 2 doubles per scalar-reducing call, the potrf status word
        scalar_calls = sum(1 for n, _, _, _ in prof if n.startswith(("gprb_lml_", "gprb_w_block_sum")))
        d2h = (16 * scalar_calls + 4 * sum(1 for n, _, _, _ in prof if n == "gprb_chol_factor")) / args.steps
        lml, grad = res
        result["e2e"] = {"value": flops / dt * 1e-9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(h2d),
                         "d2h_bytes_per_step": int(d2h), "ms_per_step": dt * 1e3,
                         "call": "GP.log_marginal_likelihood(theta, eval_gradient=True) from pinned host packed arrays",
                         "device_ms_by_entry_point": {k: round(v, 3) for k, v in sorted(parts.items())},
                         "host_ms_inside_entry_point": {k: round(v, 3) for k, v in sorted(host_parts.items())},
                         "lml": float(lml), "lml_grad": [float(g) for g in grad]}
        gp2.release_peer()
        del gp2
        gdev.clear_cache()

    # ---- predictions per second against the full training set (rank 0's share; replicas scale linearly) ----
    if not args.no_predict:
        with_io = None
        predict_parts = {}
        try:
            gp._alpha_dev = None
            K, _, _ = gp._build_K(grad=False)
            K = gp._own(K)
            gp._alpha_dev = gp._factor(K, NOISE_E, NOISE_F)
            gp._L_dev, gp._Kinv_dev = K, None            # (K is an owned copy, see below)
            gp.set_K_inv()
            n_test = 64
            tests = [a for a, _, _ in syn.structures(n_test, nrep, seed0 + 1000)]
            for a in tests[:2]:
                gp.predict_structure(a, stress=False, return_std=True, f_tol=1e-12)
            barrier()
            t0 = time.perf_counter()
            for a in tests[:8]:
                single = gp.predict_structure(a, stress=False, return_std=True, f_tol=1e-12)
            barrier()
            single_ms = (time.perf_counter() - t0) / 8 * 1e3
            gp.predict_structures(tests[:32], return_std=True, f_tol=1e-12, batch=32)
            barrier()
            _lib.PROFILE = []
            t0 = time.perf_counter()
            res = gp.predict_structures(tests, return_std=True, f_tol=1e-12, batch=32)
            barrier()
            with_io = n_test / (time.perf_counter() - t0)
            prof, _lib.PROFILE = _lib.PROFILE, None
            for n, a, b, h in prof:
                predict_parts[n] = round(predict_parts.get(n, 0.0) + a.elapsed_time(b) / (n_test / 32), 3)
            assert abs(res[7][0] - single[0]) <= 1e-8 and np.abs(res[7][1] - single[1]).max() <= 1e-8
        except Exception as exc:      # the prediction leg must not hide the covariance numbers
            result["predict_error"] = repr(exc)
        if with_io is not None:
            result["predict"] = {"value": with_io * world, "unit": "structures/s", "n_train": N, "atoms": len(tests[0]),
                                 "single_call_ms": single_ms, "device_ms_per_batch_of_32": predict_parts,
                                 "call": "GP.predict_structures(list of Atoms, return_std=True, batch=32): SO3 + K* + mean + std on "
                                         "device, host Atoms in, numpy E/F/std out; every rank predicts its own share (replicas); "
                                         "single_call_ms = one GP.predict_structure(atoms, stress=False, return_std=True)"}

    # ---- descriptor producer: SO3 of the whole training set on the device (host Atoms in, device x / dxdr / seq out) ----
    atoms_list = [a for a, _, _ in labelled]
    des.calculate_batch(atoms_list[:64], to_host=False)
    barrier()
    t0 = time.perf_counter()
    n_seq = 0
    for s0 in range(0, len(atoms_list), 64):
        r = des.calculate_batch(atoms_list[s0:s0 + 64], to_host=False)
        n_seq += int(r["seq"].shape[0])
    barrier()
    dt_so3 = time.perf_counter() - t0
    n_atoms = sum(len(a) for a in atoms_list)
    result["so3"] = {"value": len(atoms_list) / dt_so3, "unit": "structures/s", "atoms_per_structure": n_atoms // len(atoms_list),
                     "seq_rows": n_seq, "output_gb_per_s": 8.0 * (n_atoms * D + n_seq * 3 * D) / dt_so3 * 1e-9,
                     "call": "SO3.calculate_batch(64 structures, to_host=False): neighbour search, radial integrals, power spectrum "
                             "and dx/dr on device; bound by FP64 special functions, not HBM"}

    # ---- CPU baseline on the box's host cores (rank 0, N = 1 only) ------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = args.cpu_structures or 8
        gf, dt, kind, sample = run_cpu(n_cpu, nrep, seed0, 1, 0, threads)
        result["cpu_baseline"] = {"value": gf, "unit": "GFLOP/s", "cores": threads, "kind": kind, "sample": sample,
                                  "seconds": dt}
        if SO3_CPU_SECONDS:
            result["so3"]["cpu_port_structures_per_s"] = 1.0 / float(np.mean(SO3_CPU_SECONDS))
            result["so3"]["cpu_port"] = "oracle/so3.py (numpy restatement of SO3.calculate), 1 core, same structures"
    if world > 1 and getattr(gp, "_peer", None) is None and gdist.peer_gather_enabled():
        result["config"]["gather"] = "NCCL all-gather (peer mapping unavailable)"
    if rank == 0:
        print(json.dumps(result))
    if world > 1:
        gp.release_peer()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
