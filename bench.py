#!/usr/bin/env python
"""bench.py -- throughput of the tracer-advection hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--ne 120] [--qsize 35] [--test 11]

One bench "step" is one remap cycle of the reference's time loop (prim_run_subcycle, reference
src/share/prim_driver_mod.F90:701-854): rsplit=3 tracer steps (prescribed winds + 3-stage SSP-RK euler_step with
limiter 8, stage-3 hyperviscosity, 5 DSS exchanges) + 1 vertical PPM remap.  The default K=16 is the reference's
ne120 perf case (1 model-hour = 48 tracer steps, test/run_ne120_perf.sh).

  value  : tracer-steps/s = qsize * 3K / T, device-resident run (IC and winds generated on the device)
  e2e    : the same loop driven through the C ABI with HOST buffers: every tracer step the host hands derived%vn0 and
           derived%dp over from pinned memory (tse_set_derived), every cycle it reads ps_v and the tracer masses back
  roofline: the dominant kernel (k_euler_stage, 3 launches per tracer step), CUDA-event timed on the library's stream
  cpu_baseline: the CPU oracle (port of the reference; the Fortran/MPI reference cannot be built here) on the host cores
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
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

NU_Q = {8: 6e16, 30: 1e15, 120: 1e13}
TSTEP = {8: 400.0, 30: 300.0, 120: 75.0}
METRIC = "tracer-steps/sec (ne120 qsize=35 72L perf case; model-days/wall-sec and HBM GB/s in extras)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(ne, qsize, world):
    """dram__bytes_read.sum + dram__bytes_write.sum per k_euler_stage launch from the committed ncu capture of this configuration
    (profiles/traffic.json, written by tools/traffic_from_launch_list.py); None if there is no capture for it."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if world != 1 or not os.path.exists(p):
        return None
    d = json.load(open(p)).get("ne%d_q%d" % (ne, qsize))
    return d["k_euler_stage_dram_bytes_per_launch"] if d else None


def alg_bytes_per_tracer_step(nelem, qsize, rsplit=3):
    """(13 + 2/rsplit) N_q + 30 N_lev"""
    n_lev = nelem * 16 * 72 * 8
    n_q = qsize * n_lev
    return (13 + 2.0 / rsplit) * n_q + 30 * n_lev, n_q, n_lev


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for nme, val in zip(names, r[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


SAMPLE_NE = 30  # CPU sample mesh of both arms (5400 of the 86400 elements of ne120; the reference's own ne30 perf case)


def host_threads():
    """Threads the CPU arm may use: the cores this process is allowed on (torchrun exports OMP_NUM_THREADS=1, which would
    otherwise silently turn the OpenMP oracle into a one-thread run)."""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_oracle_rate(ne_target, qsize, test, steps, warmup, ne_sample=SAMPLE_NE):
    """Times the CPU oracle (OpenMP over elements, like the reference's HORIZ_OPENMP) on a bounded sample of the workload:
    the same case on the ne=30 cubed sphere; cost is linear in the element count, so the rate is scaled by the element ratio.
    Returns (tracer-steps/s at ne_target, threads used, sample description, seconds per remap cycle at ne_target)."""
    from helpers import make_oracle
    from oracle.oracle_lib import set_num_threads
    threads = set_num_threads(host_threads())
    nelem_target = 6 * ne_target * ne_target
    ne_s = min(ne_sample, ne_target)
    tstep = TSTEP.get(ne_target, 75.0)
    m, v, hv, o = make_oracle(ne_s, qsize, test, nu_q=NU_Q.get(ne_target, 1e13))
    for _ in range(warmup):
        o.prim_run_subcycle(tstep)
    t0 = time.time()
    for _ in range(steps):
        o.prim_run_subcycle(tstep)
    T = time.time() - t0
    nelem_s = 6 * ne_s * ne_s
    scale = nelem_target / nelem_s
    rate = qsize * 3 * steps / (T * scale)
    sample = "ne=%d (%d of %d elements), qsize=%d, 72L, %d remap cycle(s) = %d tracer steps + %d remap(s) after %d warm-up cycle(s); " \
             "%.2f s wall on %d OpenMP threads; rate scaled by the element ratio" % (ne_s, nelem_s, nelem_target, qsize, steps, 3 * steps,
                                                                                   steps, warmup, T, threads)
    return rate, threads, sample, T / steps * scale


def bind_near_gpu(local):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (sysfs); returns the previous affinity, or None if the
    topology is not visible (then nothing changes)."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        old = os.sched_getaffinity(0)
        cpus &= old
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return old
    except (OSError, ValueError, AttributeError):
        return None


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  The Fortran/MPI reference cannot be compiled in this
    image (no Fortran compiler), so this is the oracle port, with all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # bounded: at most 8 timed cycles and one warm-up cycle of the ne=30 sample (about 15 s each on 16 cores)
    k_eff, w_eff = max(1, min(args.steps, 8)), min(args.warmup, 1)
    rate, cores, sample, sec_per_cycle = cpu_oracle_rate(args.ne, args.qsize, args.test, k_eff, w_eff)
    if k_eff != args.steps or w_eff != args.warmup:
        sample += "; --steps %d --warmup %d were capped to %d timed + %d warm-up cycle(s)" % (args.steps, args.warmup, k_eff, w_eff)
    tstep = TSTEP.get(args.ne, 75.0)
    out = {"impl": "reference", "metric": METRIC, "value": rate, "unit": "tracer-steps/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": sec_per_cycle * 1e3, "higher_is_better": True,
           "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": workload_config(args),
           "model_days_per_wall_s": 3 * tstep / 86400.0 / sec_per_cycle,
           "cpu_baseline": {"value": rate, "unit": "tracer-steps/s", "cores": cores, "kind": "port", "sample": sample},
           "e2e": {"value": rate, "unit": "tracer-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def workload_config(args):
    return {"workload": "ne%d DCMIP1-%d perf case: qsize=%d, 72L, tstep=%gs, rsplit=3, limiter_option=8, nu_q=%g; "
                        "step = 1 remap cycle (3 tracer steps + 1 vertical_remap)" % (args.ne, args.test - 10, args.qsize,
                                                                                      TSTEP.get(args.ne, 75.0), NU_Q.get(args.ne, 1e13)),
            "ne": args.ne, "nelem": 6 * args.ne * args.ne, "qsize": args.qsize, "nlev": 72, "np": 4,
            "partition": "space-filling curve, %d rank(s)" % args.gpus,
            "l2": "inputs larger than L2 (one Qdp time level = %.2f GB per GPU)" % (6 * args.ne ** 2 * 16 * 72 * 8 * args.qsize / 1e9 / args.gpus)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from transport_se_b200.mesh import Mesh, load_vcoord
    from transport_se_b200.advection import TracerAdvection

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run)" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    ne, qsize, test = args.ne, args.qsize, args.test
    tstep, nu_q = TSTEP.get(ne, 75.0), NU_Q.get(ne, 1e13)
    mesh = Mesh(ne)
    view = mesh.local_view(rank, world)
    hv = load_vcoord()
    adv = TracerAdvection(mesh, view, hv, qsize=qsize, nu_q=nu_q, device=local)
    if world > 1:
        adv.comm_init(dist, rank, world)
    nelem_local = view.nelemd

    def barrier():
        adv.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- device-resident run (value) ----------------
    adv.dcmip_init(test)
    mass0 = adv.diag_mass(1)
    nstep = 0
    for _ in range(args.warmup):
        nstep = adv.prim_run_subcycle(tstep, nstep)
    barrier()
    adv.timer_reset()
    l0, s0 = adv.launch_count, adv.stage_launch_count
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    adv.mark(0)
    for _ in range(args.steps):
        nstep = adv.prim_run_subcycle(tstep, nstep)
    adv.mark(1)
    barrier()
    T_ms = adv.mark_elapsed_ms(0, 1)
    clk = clocks.stop() if rank == 0 else None
    launches = adv.launch_count - l0
    stage_launches = adv.stage_launch_count - s0
    stage_ms = adv.timer_ms("k_euler_stage")
    timers = {k: adv.timer_ms(k) for k in ("prim_run", "prim_advance_exp", "prim_advec_tracers_remap_rk2", "euler_step", "vertical_remap", "bndry_exchange", "edge_pack")}
    if world > 1:
        t = torch.tensor([T_ms, stage_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        T_ms, stage_ms = float(t[0]), float(t[1])
    mass = adv.diag_mass(1 if (nstep % 2 == 0) else 2)  # the fresh level after TimeLevel_update is n0_qdp (time_mod.F90:85-109)
    qmn, qmx = adv.diag_qminmax(1 if (nstep % 2 == 0) else 2)
    fhash = adv.diag_field_hash(1 if (nstep % 2 == 0) else 2)
    # size-independent check at full scale: tracer mass is conserved to roundoff (limiter, DSS, biharmonic and remap all conserve).
    # Only the 4 analytic tracers count: the checkerboard fillers (tracers 5..) start from a field that is discontinuous across
    # element edges, so their mass moves by O(1e-4) in the first DSS projections -- in the oracle by the same amount
    # (tests/test_oracle_golden.py); reported separately.
    rel = [abs(a - b) / abs(b) if b != 0.0 else 0.0 for a, b in zip(mass, mass0)]
    analytic = [0, 1, 2, 3] if test == 11 else [1]   # DCMIP 1-2: tracer 2 is the Hadley layer, all others are checkerboard fillers
    analytic = [i for i in analytic if i < qsize]
    mass_drift = float(max(rel[i] for i in analytic))
    mass_drift_fill = float(max([rel[i] for i in range(qsize) if i not in analytic] or [0.0]))
    if mass_drift > 1e-12 and not os.environ.get("TSE_BENCH_NO_CHECK"):  # (the override is for timing experiments with broken kernels)
        raise SystemExit("bench.py: tracer mass not conserved (relative drift %.3e): results are wrong" % mass_drift)

    # ---------------- end-to-end through the C ABI with host buffers (e2e) ----------------
    e2e = None
    if not args.no_e2e:
        # host copies of the prescribed winds: what a Fortran host's prim_advance_exp would hand over each step
        pin = lambda shape: torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
        # several ranks per host: allocate (first-touch) the pinned buffers on the NUMA node of the rank's GPU, so that 8 uploads of
        # 0.3 GB per step do not cross the socket interconnect
        old_aff = bind_near_gpu(local) if world > 1 else None
        vn0_h, dp_h = pin((nelem_local, 72, 2, 16)), pin((nelem_local, 72, 16))
        ps_h = pin((nelem_local, 16))
        vn0_h[:] = 0.0
        dp_h[:] = 0.0
        ps_h[:] = 0.0
        adv.get_wind(vn0_h, dp_h)
        k_e2e = max(1, min(args.steps, args.e2e_steps))
        ns = nstep

        def cycle(ns):
            # The upload of a step's winds is queued right after the previous step's kernels (tse_set_derived is asynchronous:
            # own stream, second device copy), so it overlaps them; the blocking reads come last.
            for r in range(3):
                if r > 0:
                    ns += 1
                    adv.set_derived(vn0=vn0_h, dp=dp_h)                   # H2D from pinned host memory, every tracer step
                adv.prim_advec_tracers_remap_rk2(tstep, ns)
            np1_qdp = 2 if ns % 2 == 0 else 1
            adv.vertical_remap(3 * tstep, 0, np1_qdp)
            adv.set_derived(vn0=vn0_h, dp=dp_h)                           # winds of the next cycle's first step
            adv.get_dp3d_ps(None, ps_h)                                   # D2H of the cycle's result
            adv.diag_mass(np1_qdp)
            return ns + 1
        adv.set_derived(vn0=vn0_h, dp=dp_h)
        ns = cycle(ns)  # warm-up
        barrier()
        t0 = time.perf_counter()
        adv.mark(2)
        for _ in range(k_e2e):
            ns = cycle(ns)
        adv.mark(3)
        barrier()
        wall = time.perf_counter() - t0
        e_ms = max(adv.mark_elapsed_ms(2, 3), wall * 1e3)
        if world > 1:
            t = torch.tensor([e_ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t[0])
        if old_aff is not None:
            os.sched_setaffinity(0, old_aff)
        e2e = {"value": qsize * 3 * k_e2e / (e_ms * 1e-3), "unit": "tracer-steps/s",
               "h2d_bytes_per_step": int(3 * (vn0_h.nbytes + dp_h.nbytes)), "d2h_bytes_per_step": int(ps_h.nbytes + 8 * qsize),
               "steps": k_e2e, "ms_per_step": e_ms / k_e2e,
               "note": "winds re-sent from pinned host memory every tracer step (same field each step: throughput only)"}

    if rank != 0:
        return
    # ---------------- report ----------------
    nelem = mesh.nelem
    T = T_ms * 1e-3
    n_tracer_steps = 3 * args.steps
    value = qsize * n_tracer_steps / T
    alg, n_q, n_lev = alg_bytes_per_tracer_step(nelem, qsize)
    peak, peak_src = peaks()
    # dominant kernel: k_euler_stage.  Algorithmic bytes per launch: stage 1 and 2 read + write one Qdp level (2 N_q), stage 3
    # also reads the DSS'd laplacian (3 N_q): (2+2+3)/3 N_q per launch on average, per GPU.
    stage_bytes = (7.0 / 3.0) * n_q / world
    stage_avg_s = stage_ms * 1e-3 / max(1, stage_launches)
    achieved = stage_bytes / stage_avg_s / 1e9
    out = {"metric": METRIC, "value": value, "unit": "tracer-steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": T_ms / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
           "dtype": "f64", "data": "synthetic", "config": workload_config(args),
           "model_days_per_wall_s": n_tracer_steps * tstep / 86400.0 / T,
           "ms_per_tracer_step": T_ms / n_tracer_steps,
           "step_hbm": {"alg_bytes_per_tracer_step": alg, "achieved_gbs_per_gpu": alg * n_tracer_steps / T / 1e9 / world, "peak_gbs": peak,
                        "frac": alg * n_tracer_steps / T / 1e9 / world / peak, "frac_of_8TBs": alg * n_tracer_steps / T / 1e9 / world / 8000.0},
           "roofline": {"bound": "hbm", "kernel": "k_euler_stage", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": measured_traffic(ne, qsize, world), "peak_source": peak_src, "alg_bytes_per_launch": stage_bytes,
                        "avg_launch_ms": stage_avg_s * 1e3, "launches": int(stage_launches),
                        "share_of_step": stage_ms / T_ms},
           "timers_ms": timers, "gpu_launches": int(launches), "clocks": clk, "e2e": e2e,
           "tracer_mass": [float(x) for x in mass[:4]],
           # order-independent sums and extrema: bitwise equal for any --gpus N when the fields are (compare the lines of a scaling run)
           # per-tracer fingerprint of the whole final Qdp field (wrapping sum of mix(bits, global position), integer all-reduce):
           # equal values for any --gpus N <=> the fields are bit-for-bit equal (README:46-47); field_hash_all folds all tracers
           "field_hash_hex": ["%016x" % int(x) for x in fhash[:6]],
           "field_hash_all": "%016x" % (sum(int(x) * (2 * i + 1) for i, x in enumerate(fhash)) % (1 << 64)),
           "tracer_mass_hex": [float(x).hex() for x in mass[:6]], "tracer_qmin_qmax_hex": [[float(a).hex(), float(b).hex()] for a, b in zip(qmn[:6], qmx[:6])], "mass_drift_rel": mass_drift, "mass_drift_rel_checkerboard": mass_drift_fill, "device_bytes": int(adv.device_bytes),
           "published_context": "reference Fortran/MPI on 960 Edison cores: 42.6 s per model-hour = 39.4 tracer-steps/s (README:174)"}
    if not args.no_cpu and world == 1:
        # same sample and protocol as --impl reference, shortened: one warm-up cycle (first-touch of the oracle's arrays), one timed
        rate, cores, sample, _ = cpu_oracle_rate(ne, qsize, test, 1, 1)
        out["cpu_baseline"] = {"value": rate, "unit": "tracer-steps/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(out))
    adv.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    # libraries (NCCL, torchrun) may write to stdout: keep fd 1 for the single JSON line, send everything else to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ne", type=int, default=120)
    ap.add_argument("--qsize", type=int, default=35)
    ap.add_argument("--test", type=int, default=11)
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    sys.stdout.flush()


if __name__ == "__main__":
    main()
