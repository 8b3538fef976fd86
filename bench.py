#!/usr/bin/env python
"""Benchmark of the hot path: batched real-time-iteration NMPC steps for the Crazyflie OCP.

    python bench.py --gpus N --steps K --warmup W            # our CUDA path (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

A step = one RTI solve (preparation + feedback = one acados_solve()) of every instance of the batch from the
same seeded initial iterate.  Workload = BASELINE.json configs[1] per GPU: 65,536 hover-regulation OCPs,
N=50, nx=13, nu=4, random feasible x0 (crazyflie_nmpc_b200.workloads.hover_batch).  Weak scaling: the per-GPU
batch is fixed, ranks are independent (no collective inside the solve); with N>1 ranks the solved first
controls u0 are all-gathered over NCCL after each step (BASELINE config 4), inside the timed region.

Prints ONE JSON line (rank 0). Contract keys: see the task description; extra objects: roofline, cpu_baseline.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "NMPC solves/sec (batch, device-timed) N=50 nx=13 nu=4"
UNIT = "solves/s"
TS = 0.015


def alg_bytes(N):
    """Algorithmic (compulsory) bytes per solve, SURVEY.md 8(d): read x0, read yref, read + write the iterate."""
    return 8 * (51 * N + 52)


def make_workload(name, B, N, seed):
    from crazyflie_nmpc_b200 import workloads as wl
    if name == "helix":
        return wl.helix_batch(B, N, seed=seed)
    return wl.hover_batch(B, N, seed=seed)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"   # /opt/skills/guides/B200_PROFILING.md fallback


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        time.sleep(0.25)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, reasons = [], set()
        for r in rows:
            try:
                r = [c.strip() for c in r]
                sm.append(float(r[1]))
                out["sm_max_mhz"] = float(r[2])
                for name, c in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if c.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if sm:
            out["sm_mhz"] = float(np.median(sm))
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------ reference / CPU arm
def cpu_solver():
    """(callable, kind): the reference's own acados/HPIPM/BLASFEO build when it travelled with the repo
    (oracle/_ref), else the plain-C port oracle."""
    from oracle import oracle as orc
    if orc.ref_available():
        ref = orc.Ref()

        def run(N, w, nthreads, cond_N=0):
            x, u = w["x_init"].copy(), w["u_init"].copy()
            st, it, tt = ref.batch(N, TS, w["x0"], w["yref"], w["yref_e"], x, u, nthreads=nthreads, cond_N=cond_N)
            return x, u, st, float(tt.sum())
        return run, "reference"
    port = orc.Port()

    def run(N, w, nthreads, cond_N=0):
        x, u = w["x_init"].copy(), w["u_init"].copy()
        st, it = port.batch(N, TS, w["x0"], w["yref"], w["yref_e"], x, u)
        return x, u, st, float("nan")
    return run, "port"


def time_cpu(N, workload, sample, nthreads, seed):
    run, kind = cpu_solver()
    if kind == "port":
        nthreads = 1
    w = make_workload(workload, sample, N, seed)
    t0 = time.perf_counter()
    _, _, st, t_own = run(N, w, nthreads)
    dt = time.perf_counter() - t0
    out = dict(value=sample / dt, unit=UNIT, cores=nthreads, kind=kind,
               sample=f"{sample} seeded instances of the same workload, one RTI step each, wall clock over {nthreads} host threads"
                      + (f"; acados' own time_tot sums to {t_own / sample * 1e3:.3f} ms/solve/core" if t_own == t_own else ""),
               failures=int((st != 0).sum()))
    if kind == "reference" and N % 10 == 0:
        # fair to the reference: the same solver with real partial condensing, which the reference's own
        # configuration (qp_cond_N = N) does not use; `value` above stays the reference's configuration
        best = None
        for cn in (N // 5, N // 10):
            t0 = time.perf_counter()
            run(N, w, nthreads, cond_N=cn)
            v = sample / (time.perf_counter() - t0)
            if best is None or v > best[1]:
                best = (cn, v)
        out["with_partial_condensing"] = dict(qp_cond_N=best[0], value=best[1], unit=UNIT)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    sample = args.ref_sample or 192 * cores
    vals = []
    for i in range(args.warmup + args.steps):
        r = time_cpu(args.horizon, args.workload, sample, cores, seed=args.seed + i)
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals])) if vals else 0.0
    cb = dict(vals[-1]) if vals else {}
    cb["value"] = v
    line = dict(impl="reference", metric=METRIC, value=v, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=(sample / v * 1e3) if v else None, higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f64", data="synthetic",
                config=dict(workload=f"BASELINE configs[{2 if args.workload == 'helix' else 1}]: batch={args.batch} {args.workload} "
                                     f"OCPs per GPU, N={args.horizon}, nx=13, nu=4, random feasible x0, 1 RTI step per step",
                            sample=f"bounded sample of {sample} seeded instances of that workload per step (CPU throughput per solve does "
                                   f"not depend on the batch size)", batch_per_step=sample, horizon=args.horizon, host_threads=cores,
                            qp_cond_N=args.horizon),
                value_best_config=(dict(cb["with_partial_condensing"], note="the same reference solver with real partial condensing "
                                        "(not its own configuration, which is qp_cond_N = N)") if "with_partial_condensing" in cb else None),
                cpu_baseline=cb, e2e=dict(value=v, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0), gpu_launches=0)
    print(json.dumps(line))


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import crazyflie_nmpc_b200 as cf
    from crazyflie_nmpc_b200 import sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the solve path has no CPU implementation")
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]        # those levels print "NCCL version ..." on stdout: keep it to the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, B = args.horizon, args.batch
    w = make_workload(args.workload, B, N, args.seed + rank)
    # a dedicated torch stream carries every copy, kernel and event of the benchmark
    stream = torch.cuda.Stream(device=local)
    torch.cuda.set_stream(stream)
    s = cf.BatchSolver(B, N, TS, device=local)
    s.set_stream(stream.cuda_stream)
    if args.qp_cond_N:
        s.set_option("qp_cond_N", args.qp_cond_N)

    dev = torch.device("cuda", local)
    d_in = {k: torch.from_numpy(w[k]).to(dev) for k in ("x0", "yref", "yref_e", "x_init", "u_init")}
    pin = {k: torch.from_numpy(w[k]).pin_memory() for k in ("x0", "yref", "yref_e")}
    u0_dev = torch.empty(B, 4, dtype=torch.float64, device=dev)
    u0_host = torch.empty(B, 4, dtype=torch.float64).pin_memory()
    st_host = torch.empty(B, dtype=torch.int32).pin_memory()
    s.set("x0", d_in["x0"]).set("yref", d_in["yref"]).set("yref_e", d_in["yref_e"])

    def step_device():
        # inputs resident in HBM: restore the seeded initial iterate (device copy), solve, gather u0
        s.set("x", d_in["x_init"]).set("u", d_in["u_init"])
        s.solve(1)
        if world > 1:
            s.get("u", 0, out=u0_dev)
            sharding.gather_u0(u0_dev, world * B)

    def step_e2e():
        # what a caller of the reference API does per tick, with host buffers: x0 + yref in, u0 + status out
        s.set("x", d_in["x_init"]).set("u", d_in["u_init"])
        s.solve_from_host(pin["x0"], pin["yref"], pin["yref_e"], n_chunks=args.e2e_chunks)   # upload overlapped with the solve
        s.get("u", 0, out=u0_dev)
        u0_host.copy_(u0_dev, non_blocking=True)
        s.get("status", 0, out=st_host)   # host destination: synchronises the stream
        if world > 1:
            sharding.gather_u0(u0_dev, world * B)

    # the same tick through the closed-loop driver: only the measured states travel to the device, the
    # reference window is built there from (policy, trajectory row | set-point); motor commands and twist travel back
    from crazyflie_nmpc_b200 import workloads as wl
    motors_host = torch.empty(B, 4, dtype=torch.int32).pin_memory()
    twist_host = torch.empty(B, 4, dtype=torch.float64).pin_memory()
    if args.workload == "helix":
        s.set_trajectory(wl.helix_table())
        d_pol = torch.full((B,), cf.POLICY_TRACKING, dtype=torch.int32, device=dev)
        d_it = torch.from_numpy(np.ascontiguousarray(w["i0"], dtype=np.int32)).to(dev)
    else:
        d_pol = torch.full((B,), cf.POLICY_REGULATION, dtype=torch.int32, device=dev)
        d_it = torch.zeros(B, dtype=torch.int32, device=dev)
        s.set("setpoint", np.ascontiguousarray(w["yref"][:, 0, :3]))
    s.set("uss", [wl.hover_speed() if args.workload == "hover" else 15.7777])

    def step_tick():
        s.set("x0", pin["x0"])
        s.set("policy", d_pol).set("traj_iter", d_it)            # same window every step (device-side, 0.5 MB)
        s.set("x", d_in["x_init"]).set("u", d_in["u_init"])
        s.tick()
        s.get("motors", out=motors_host)
        s.get("twist", out=twist_host)                           # host destination: synchronises the stream
        if world > 1:
            s.get("u", 0, out=u0_dev)
            sharding.gather_u0(u0_dev, world * B)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, per_step=None):
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        l0 = s.info("launches") if s is not None else 0
        ev[0].record(stream)
        for i in range(steps):
            fn()
            ev[i + 1].record(stream)
        barrier()
        ms = ev[0].elapsed_time(ev[steps])
        if per_step is not None:
            per_step.extend(ev[i].elapsed_time(ev[i + 1]) for i in range(steps))
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, s.info("launches") - l0

    for _ in range(max(args.warmup, 3)):
        step_device()
    # kernel-only time of the dominant kernel, CUDA events on its own stream, averaged over the timed steps
    sampler = ClockSampler(local) if rank == 0 else None
    step_ms = []
    ms_total, launches = timed(step_device, args.steps, step_ms)
    kern_ms, phase_ms = [], []
    for _ in range(min(args.steps, 5)):
        s.set("x", d_in["x_init"]).set("u", d_in["u_init"])
        s.solve(1)
        kern_ms.append(s.last_solve_ms())
        phase_ms.append(s.last_phase_ms())
    clocks = sampler.stop() if sampler else None
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    u0_e2e = u0_host.clone()
    for _ in range(2):
        step_tick()
    ms_tick, _ = timed(step_tick, args.steps)
    # the driver built the same references: same first controls as the host-fed path (motors = trunc(u0))
    tick_consistent = bool((motors_host.numpy() == u0_e2e.numpy().astype(np.int32)).all())

    # the same device-timed step with the reference's per-iteration linear-system residual checks switched on
    # (x_ocp_qp_ipm.c:2029-2059,2311-2318: what feeds its LQ / iterative-refinement nets; flags only, same results)
    ms_chk = None
    if not args.qp_cond_N:
        s.set_option("lin_res_check", 1)
        step_device()
        ms_chk, _ = timed(step_device, max(2, args.steps // 2))
        ms_chk /= max(2, args.steps // 2)
        flags_chk = int((s.get("flags") & 3 != 0).sum())
        s.set_option("lin_res_check", 0)

    # host -> device rate of this rank's pinned input upload, all ranks copying at the same time (names the e2e limiter)
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record(stream)
    for _ in range(3):
        d_in["yref"].copy_(pin["yref"], non_blocking=True)
    h1.record(stream)
    barrier()
    h2d_gbs = 3 * pin["yref"].numel() * 8 / (h0.elapsed_time(h1) * 1e-3) / 1e9
    if world > 1:
        t = torch.tensor([h2d_gbs], dtype=torch.float64, device=dev)
        tl = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(tl, t)
        h2d_all = [float(v.item()) for v in tl]
    else:
        h2d_all = [h2d_gbs]

    # N > 1: the gathered first controls must contain every rank's own results, and a sample of each rank's solves is
    # checked against the CPU oracle (checker only: outside every timed region)
    multi = None
    if world > 1:
        step_device()
        s.get("u", 0, out=u0_dev)
        allu = sharding.gather_u0(u0_dev, world * B)
        own_ok = bool(torch.equal(allu[rank * B:(rank + 1) * B], u0_dev))
        from oracle.oracle import Port
        idx = np.arange(0, B, B // 16)[:16]
        xo, uo = np.ascontiguousarray(w["x_init"][idx]), np.ascontiguousarray(w["u_init"][idx])
        Port().batch(N, TS, np.ascontiguousarray(w["x0"][idx]), np.ascontiguousarray(w["yref"][idx]),
                     np.ascontiguousarray(w["yref_e"][idx]), xo, uo, cond_N=args.qp_cond_N)
        ug = s.get("u_all")[idx]
        err = float((np.abs(ug - uo) / (1 + np.abs(uo))).max())
        t = torch.tensor([0.0 if own_ok else 1.0, err], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        multi = dict(gathered_u0_contains_every_rank=bool(t[0].item() == 0.0), oracle_checked_per_rank=16,
                     max_rel_err_vs_oracle=float(t[1].item()), tolerance=1e-6)
        if not multi["gathered_u0_contains_every_rank"] or multi["max_rel_err_vs_oracle"] > 1e-6:
            raise SystemExit(f"bench.py: multi-GPU result check failed: {multi}")

    # BASELINE configs[3] (262,144 helix OCPs over 8 GPUs = 32,768 per GPU, NCCL all-gather of u0): measured inside the
    # multi-GPU run, with and without the gather
    config4 = None
    if world > 1 and not args.no_config4:
        B4 = 262144 // 8
        w4 = make_workload("helix", B4, N, args.seed + 1 + rank)
        s4 = cf.BatchSolver(B4, N, TS, device=local)
        s4.set_stream(stream.cuda_stream)
        if args.qp_cond_N:
            s4.set_option("qp_cond_N", args.qp_cond_N)
        d4 = {k: torch.from_numpy(w4[k]).to(dev) for k in ("x0", "yref", "yref_e", "x_init", "u_init")}
        s4.set("x0", d4["x0"]).set("yref", d4["yref"]).set("yref_e", d4["yref_e"])
        u4 = torch.empty(B4, 4, dtype=torch.float64, device=dev)

        def step4(gather):
            s4.set("x", d4["x_init"]).set("u", d4["u_init"])
            s4.solve(1)
            s4.get("u", 0, out=u4)
            if gather:
                sharding.gather_u0(u4, world * B4)
        for _ in range(3):
            step4(True)
        ms_g, _ = timed(lambda: step4(True), args.steps)
        ms_n, _ = timed(lambda: step4(False), args.steps)
        ok4 = torch.tensor([int((s4.get("status") == 0).sum())], dtype=torch.int64, device=dev)
        dist.all_reduce(ok4)
        config4 = dict(workload=f"BASELINE configs[3]: batch={world * B4} helix OCPs over {world} GPUs ({B4} per GPU), NCCL all-gather of u0",
                       value_with_gather=world * B4 / (ms_g / args.steps * 1e-3), value_without_gather=world * B4 / (ms_n / args.steps * 1e-3),
                       unit=UNIT, ms_per_step_with_gather=ms_g / args.steps, ms_per_step_without_gather=ms_n / args.steps,
                       status_ok=int(ok4.item()), of=world * B4)
        s4.close()

    # sanity: the timed work really solved the batch
    status = s.get("status")
    iters = s.get("qp_iter")
    ok = int((status == 0).sum())
    if world > 1:
        t = torch.tensor([ok], dtype=torch.int64, device=dev)
        dist.all_reduce(t)
        ok = int(t.item())

    if rank == 0:
        ms_step = ms_total / args.steps
        value = world * B / (ms_step * 1e-3)
        e2e_v = world * B / (ms_e2e / args.steps * 1e-3)
        peaks, how = measured_peaks()
        # the RTI step is two launches by default (preparation kernel, feedback kernel): the algorithmic bytes belong to
        # the step, so they are set against the sum; the feedback kernel is the dominant one
        k_ms = float(np.mean(kern_ms))
        prep_ms, fb_ms = (float(v) for v in np.mean(np.array(phase_ms), axis=0))
        two = bool(s.info("two_kernels"))
        achieved = B * alg_bytes(N) / (k_ms * 1e-3) / 1e9
        traffic = None
        tp = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get(f"bytes_per_launch_B{B}_N{N}")
            except Exception:
                traffic = None
        it_mean = float(iters.mean())
        # secondary figure of SURVEY 8(d): fp64 FLOPs of one solve, N (13.5k + n_ipm (9.5k + 1.5 * 1.5k + 3 * 1.5k)),
        # against the measured fp64 FMA peak of this device (cfnmpc_measure_fp64_peak)
        flop_per_solve = N * (13500.0 + it_mean * (9500.0 + 1.5 * 1500.0 + 3 * 1500.0))
        fp64_peak = cf.measure_fp64_peak(local)
        tf = B * flop_per_solve / (k_ms * 1e-3) / 1e12
        pc = bool(args.qp_cond_N)
        kname = (f"cf_pcond_kernel<BS={s.info('pcond_block_size')}> (dominant; qp_cond_N={args.qp_cond_N}) after cf_rti_kernel<4,3,PREPARATION>" if pc else
                 "cf_rti_kernel<4,4,FEEDBACK> (dominant) after cf_rti_kernel<4,3,PREPARATION>" if two else "cf_rti_kernel<4,3> (fused)")
        roofline = dict(bound="hbm", achieved=achieved, peak=peaks["hbm_gbs"], unit="GB/s", frac=achieved / peaks["hbm_gbs"],
                        traffic=None if pc else traffic, kernel_ms=k_ms, dominant_kernel_ms=fb_ms if two else k_ms, peak_source=how,
                        kernel=kname,
                        kernels_ms=dict(preparation=prep_ms, feedback=fb_ms) if two else None,
                        alg_bytes_per_solve=alg_bytes(N),
                        achieved_dominant_kernel=B * alg_bytes(N) / ((fb_ms if two else k_ms) * 1e-3) / 1e9,
                        flops=dict(mflop_per_solve=flop_per_solve / 1e6, achieved_tflops=tf, peak_tflops=fp64_peak,
                                   frac=tf / fp64_peak if fp64_peak else None, peak_source="measured: cfnmpc_measure_fp64_peak (DFMA stream)"),
                        note="algorithmic bytes = x0 + yref + iterate in/out (SURVEY 8d); real traffic is dominated by factor/linearisation spill")
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            cpu = time_cpu(N, args.workload, args.ref_sample or min(B, 512 * cores), cores, args.seed)
        h2d = sum(pin[k].numel() * 8 for k in pin)
        d2h = u0_host.numel() * 8 + st_host.numel() * 4
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
                    ms_per_step=ms_step, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f64", data="synthetic",
                    config=dict(workload=f"BASELINE configs[{2 if args.workload == 'helix' else 1}]: batch={B} {args.workload} "
                                         f"OCPs per GPU, N={N}, nx=13, nu=4, random feasible x0, 1 RTI step per step",
                                batch_per_gpu=B, global_batch=world * B, horizon=N, qp_cond_N=args.qp_cond_N or N, parallelism=f"dp{world} (independent shards"
                                + (", NCCL all-gather of u0)" if world > 1 else ")"),
                                l2="inputs larger than L2 (iterate+yref 900 MB, scratch %d MB per GPU)" % (s.info("scratch_bytes") >> 20),
                                occupancy=(dict(kernels="preparation + feedback", warps_per_sm=s.info("feedback_blocks_per_sm") * s.info("feedback_warps_per_block"),
                                                regs=s.info("feedback_regs_per_thread"), grid=s.info("feedback_grid"),
                                                preparation=dict(warps_per_sm=12, grid=s.info("preparation_grid")),
                                                prepared_mb=s.info("prepared_bytes") >> 20) if two else
                                           dict(kernels="fused", warps_per_sm=s.info("blocks_per_sm") * s.info("warps_per_block"),
                                                regs=s.info("regs_per_thread"), grid=s.info("grid")))),
                    clocks=clocks, e2e=dict(value=e2e_v, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h,
                                            ms_per_step=ms_e2e / args.steps, upload_chunks=args.e2e_chunks),
                    e2e_closed_loop=dict(value=world * B / (ms_tick / args.steps * 1e-3), unit=UNIT, ms_per_step=ms_tick / args.steps,
                                         h2d_bytes_per_step=pin["x0"].numel() * 8,
                                         d2h_bytes_per_step=motors_host.numel() * 4 + twist_host.numel() * 8,
                                         consistent_with_e2e=tick_consistent,
                                         note="cfnmpc_batch_tick: reference window from (policy, set-point | trajectory row) on the device, "
                                              "RTI step, motor/twist commands; only x0 goes up"),
                    ms_per_step_best=float(np.min(step_ms)), ms_per_step_median=float(np.median(step_ms)),
                    value_best=world * B / (float(np.min(step_ms)) * 1e-3) if world == 1 else None,
                    value_median=world * B / (float(np.median(step_ms)) * 1e-3) if world == 1 else None,
                    value_with_lin_res_check=(dict(value=world * B / (ms_chk * 1e-3), unit=UNIT, ms_per_step=ms_chk, flagged_instances=flags_chk,
                                                   note="same step with the reference's linear-system residual checks evaluated every "
                                                        "IPM iteration (option lin_res_check: the general kernel variants, which carry that code, as two launches); default off, always-on detection = status / "
                                                        "qp_status / BAD_PIVOT / NONFINITE flags") if ms_chk else None),
                    h2d_gbs_per_rank=dict(min=min(h2d_all), mean=float(np.mean(h2d_all)), max=max(h2d_all),
                                          note="pinned host -> device copy of this rank's yref (%.0f MB), all ranks at once" % (pin["yref"].numel() * 8 / 1e6)),
                    multi_gpu_check=multi, config4=config4,
                    gpu_launches=int(launches), roofline=roofline, cpu_baseline=cpu,
                    solved=dict(status_ok=ok, of=world * B, ipm_iter_mean=float(iters.mean()), ipm_iter_max=int(iters.max())))
        print(json.dumps(line))
    s.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="hover", choices=["hover", "helix"])
    ap.add_argument("--batch", type=int, default=65536, help="instances per GPU")
    ap.add_argument("--horizon", type=int, default=50)
    ap.add_argument("--seed", type=int, default=20261017)
    ap.add_argument("--ref-sample", type=int, default=0, help="instances per CPU step (default: scaled to the core count)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config4", action="store_true", help="skip the BASELINE configs[3] record of a multi-GPU run")
    ap.add_argument("--qp-cond-N", type=int, default=0, help="partial condensing to this many stages (0 = the reference's configuration, qp_cond_N = N)")
    ap.add_argument("--e2e-chunks", type=int, default=16, help="chunks of the overlapped host-to-device upload in the e2e path")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
