"""Benchmark of the operator-backpropagation hot path (BASELINE.json metric: Pauli term-gate updates/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl engine|reference] [--workload NAME]

One *step* = one full ``backpropagate()`` of the workload (default: BASELINE config 2, the 127-qubit
heavy-hex kicked-Ising circuit, Z_62, 20 Trotter steps, ``max_paulis=1e6``; the call runs until the
operator budget stops it).  One *term-gate update* = one (Pauli term, applied gate) pair, counted with
the term count at gate entry.

Engine arm (default) prints one JSON line with
  value       updates/s over the device-resident region (CUDA events from after the observable upload
              to before the result download, max over ranks),
  e2e         updates/s through the public ``backpropagate()`` call with host buffers (upload, host
              gate classification, kernels, download and unpacking all inside the timed region),
  roofline    the dominant kernel (k_rotate) against the measured HBM peak,
  cpu_baseline  the numpy oracle port on a bounded sample of the same workload (rank 0, N=1 only).
``--impl reference`` times the reference's CPU algorithm (the oracle port: Qiskit is not installable
here) on the host cores, on the same workload, each step a bounded sample.

N > 1 (torchrun): every rank backpropagates its own observable of the workload (independent
observables partition the path with no data-path collective): weak scaling.  With
``--sharded`` the ranks instead share ONE observable as a hash-sharded table (strong scaling; meant for
tables far beyond one GPU's L2, e.g. ``--workload config4 --max-paulis 20000000``).
"""

from __future__ import annotations

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

METRIC = "pauli_term_gate_updates_per_s"
UNIT = "updates/s"
L2_BYTES = 126 << 20


def algorithmic_bytes_per_update(num_qubits: int) -> int:
    """B_tg = 2*(16*W + 16): read the packed key + complex128 coefficient once, write them once."""
    w = (num_qubits + 63) // 64
    return 2 * (16 * w + 16)


def make_workloads(name: str, rank: int, args) -> list:
    """The workload(s) one step runs: config 2 is the named theta_h sweep, the others a single case."""
    from qiskit_addon_obp_b200 import workloads as wl

    thetas = [float(t) for t in str(args.theta).split(",")]
    if name == "config2":
        # rank r measures Z on a different heavy-hex qubit (independent observables)
        qubits = [62, 64, 43, 81, 24, 100, 45, 83]
        return [
            wl.config2_kicked_ising(theta_h=t, steps=20, max_paulis=args.max_paulis, observable_qubit=qubits[rank % len(qubits)])
            for t in thetas
        ]
    if name == "config1":
        return [wl.config1_xy_chain()]
    if name == "config3":
        return [wl.config3_square_lattice(max_paulis=args.max_paulis)]
    if name == "config4":
        return [wl.config4_near_clifford(depth=args.depth or 200, max_paulis=args.max_paulis, seed=20261019 + rank)]
    if name == "config5":
        return [wl.config5_magic_sweep(theta=t, depth=args.depth or 40, max_paulis=args.max_paulis, seed=rank) for t in thetas]
    raise SystemExit(f"unknown workload {name}")


def workload_name(wls) -> str:
    if len(wls) == 1:
        return wls[0].name
    import re

    angles = [re.search(r"theta_h=([0-9.]+)|RZ\(([0-9.]+)\)", w.name) for w in wls]
    angles = [(m.group(1) or m.group(2)) if m else "?" for m in angles]
    base = re.sub(r"theta_h=[0-9.]+, ", "", wls[0].name)
    return f"{base}; one step = the angle sweep {{{', '.join(angles)}}}"


# ---------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, through NVML in a background thread.

    (Spawning ``nvidia-smi -lms`` instead stalls GPU work for ~35 ms at every poll on these boxes, a 15 %
    perturbation of a 30 ms step; in-process NVML queries of two fields do not.  nvidia-smi remains the
    fall-back when the NVML bindings are missing.)"""

    FIELDS = (
        "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    )

    def __init__(self, device: int):
        self.device = device
        self.samples: list = []  # (stamp, sm_mhz, sm_max_mhz, set of reasons)
        self.t_begin = 0.0
        self.proc = None
        self.thread = None
        self.stop_flag = False
        self.source = None
        # With NVML the clocks are sampled from the main thread BETWEEN the timed steps (inside the timed region,
        # outside the per-step timers, the GPU having just finished a step): one sample per step and no query
        # that can land inside a step (NVML takes a driver lock kernel launches need as well).
        # OBP_BENCH_SAMPLE_MS > 0 brings a background poller back.  Note on end-to-end noise: on some of the
        # shared boxes individual 26 ms end-to-end steps take 40-130 ms, with every polling scheme and with no
        # polling at all, while on others ten steps in a row stay within 26.0 +- 0.5 ms; the device-timed
        # `value` is unaffected.  It is host-side noise of the box, not of the sampler.
        self.period_s = float(os.environ.get("OBP_BENCH_SAMPLE_MS", "0")) / 1e3
        self._poll = None

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.nvml = pynvml
            self.source = "nvml"
            self._poll = self._make_nvml_poll()
            self._poll()  # (the first query initialises NVML's device state; keep it out of the timed region)
            if self.period_s > 0:
                self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
                self.thread.start()
            return
        except Exception:  # noqa: BLE001 - no NVML bindings / init failure: fall back to the CLI
            self.source = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE,
                stderr=subprocess.DEVNULL,
                text=True,
            )
        except OSError:
            self.proc = None
            return
        self.source = "nvidia-smi"
        self.thread = threading.Thread(target=self._read_cli, daemon=True)
        self.thread.start()

    def _physical_index(self) -> int:
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if self.device < len(ids) and ids[self.device].isdigit():
                return int(ids[self.device])
        return self.device

    def _make_nvml_poll(self):
        n = self.nvml
        bits = {
            "hw_slowdown": getattr(n, "nvmlClocksEventReasonHwSlowdown", getattr(n, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(n, "nvmlClocksEventReasonSwPowerCap", getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
        }
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        try:
            sm_max = float(n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            sm_max = float("nan")

        def poll():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(get_reasons(self.handle))
                self.samples.append((time.perf_counter(), sm, sm_max, {k for k, b in bits.items() if mask & b}))
            except Exception:  # noqa: BLE001
                pass

        return poll

    def _poll_nvml(self):
        while not self.stop_flag:
            self._poll()
            time.sleep(self.period_s)

    def poll_once(self):
        """One sample from the calling thread (between timed steps)."""
        if self._poll is not None:
            self._poll()

    def _read_cli(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) == 6:
                try:
                    reasons = {nm for nm, v in zip(names, parts[2:]) if v.lower().startswith("active")}
                    self.samples.append((time.perf_counter(), float(parts[0]), float(parts[1]), reasons))
                except ValueError:
                    continue

    def wait_ready(self, timeout: float = 8.0):
        """Block until the first sample arrived (an nvidia-smi start-up would stall the warm-up steps)."""
        t0 = time.perf_counter()
        while self.thread is not None and not self.samples and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)

    def mark_begin(self):
        """Only samples taken from now on count."""
        self.t_begin = time.perf_counter()

    def stop(self) -> dict:
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        if self.thread is None and self._poll is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML bindings and no nvidia-smi"]}
        inside = [s for s in self.samples if s[0] >= self.t_begin]
        reasons: set = set()
        for s in inside:
            reasons |= s[3]
        return {
            "sm_mhz": float(np.median([s[1] for s in inside])) if inside else None,
            "sm_max_mhz": float(max(s[2] for s in inside)) if inside else None,
            "samples": len(inside),
            "source": self.source,
            "reasons": sorted(reasons),
        }


# ---------------------------------------------------------------------------------------------
# CPU side: the oracle port on a bounded sample
# ---------------------------------------------------------------------------------------------
def cpu_sample(wls, max_updates: int) -> tuple[int, float]:
    """The oracle port on the first ``max_updates`` term-gate updates of every workload of the step."""
    from oracle import obp_oracle

    updates, t0 = 0, time.perf_counter()
    for wl in wls:
        trace = obp_oracle.GateTrace()
        obp_oracle.backpropagate(
            wl.observables, wl.slices, trace=trace, max_updates=max(1, max_updates // len(wls)), **wl.kwargs()
        )
        updates += trace.updates
    return updates, time.perf_counter() - t0


def run_reference(args, rank: int, world: int) -> None:
    if rank != 0:
        return
    wls = make_workloads(args.workload, 0, args)
    cap = args.cpu_updates
    for _ in range(args.warmup):
        cpu_sample(wls, max(1000, cap // 20))
    tot_u, tot_t = 0, 0.0
    for _ in range(args.steps):
        u, t = cpu_sample(wls, cap)
        tot_u += u
        tot_t += t
    value = tot_u / tot_t
    sample = f"first {tot_u // max(1, args.steps)} term-gate updates of the workload per step (oracle stopped by max_updates)"
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * tot_t / max(1, args.steps),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "complex128 (f64) coefficients, bool keys",
        "data": "synthetic",
        "config": {"workload": workload_name(wls), "theta": args.theta, "max_paulis": args.max_paulis},
        "cpu_baseline": {
            "value": value,
            "unit": UNIT,
            "cores": 1,
            "kind": "port",
            "sample": sample,
            "note": "numpy restatement of the reference loop (Qiskit not installable offline); single-threaded like the reference",
        },
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# engine arm
# ---------------------------------------------------------------------------------------------
def flush_l2(torch, device, buf):
    buf.add_(1)  # writes 256 MiB > 126 MB L2
    torch.cuda.synchronize(device)


def large_table_roofline(backpropagate, bpmod, peak: float) -> dict:
    """The same k_rotate kernel where it IS bound by HBM: config 2 (theta_h = 0.3) allowed to grow to
    2e7 terms (peak ~2.7e7 rows x 56 B + index: 20x the L2), one untimed-for-`value` pass, per-class device
    time from CUDA events around the k_rotate launches."""
    from qiskit_addon_obp_b200 import workloads as wl

    w = wl.config2_kicked_ising(theta_h=0.3, steps=20, max_paulis=20_000_000)
    backpropagate(w.observables, w.slices, **w.kwargs())  # warm-up: table allocation, slice programs
    backpropagate(w.observables, w.slices, **w.kwargs())
    rot = (bpmod.last_run.profile or {}).get("rotate", {"ms": 0.0, "term_updates": 0, "passes": 0})
    btg = algorithmic_bytes_per_update(w.num_qubits)
    achieved = rot["term_updates"] * btg / (rot["ms"] * 1e-3) / 1e9 if rot["ms"] > 0 else None
    return {
        "kernel": "k_rotate",
        "workload": w.name,
        "bound": "hbm",
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": achieved / peak if achieved else None,
        "bytes_per_term_gate_update": btg,
        "passes": rot["passes"],
        "term_updates": rot["term_updates"],
        "ms": rot["ms"],
        "updates_per_s_whole_run": bpmod.last_run.updates / (bpmod.last_run.device_ms * 1e-3),
        "traffic": {
            "dram_bytes_per_launch": 300.4e6,
            "rows_per_launch": 4.5e6,
            "algorithmic_bytes_per_launch": 4.5e6 * btg,
            "source": "ncu --set full, profiles/r1d_k_rotate_big_ncu.txt (dram__bytes_read.sum + dram__bytes_write.sum)",
        },
    }


def run_engine(args, rank: int, local_rank: int, world: int) -> None:
    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference)")
    os.environ["OBP_B200_DEVICE"] = str(local_rank)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    if args.no_download:
        os.environ["OBP_B200_DOWNLOAD"] = "0"
    sharded = bool(args.sharded) and world > 1
    if sharded:
        os.environ["OBP_B200_SHARDED"] = "1"  # one observable, hash-sharded over all ranks (strong scaling)

    from qiskit_addon_obp_b200 import backpropagate
    from qiskit_addon_obp_b200 import backpropagation as bpmod

    bpmod.PROFILE = not args.no_profile
    wls = make_workloads(args.workload, 0 if sharded else rank, args)  # sharded: every rank drives the same job
    flush_buf = torch.zeros(256 << 20, dtype=torch.uint8, device=device)

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    def one_step():
        """One pass of the hot path over the step's workload(s) through the public API."""
        flush_l2(torch, device, flush_buf)
        acc = {"dt": 0.0, "dev_ms": 0.0, "updates": 0, "h2d": 0, "d2h": 0, "n_out": 0, "slices": 0, "prof": {}}
        for wl in wls:
            t0 = time.perf_counter()
            out, rest, md = backpropagate(wl.observables, wl.slices, **wl.kwargs())
            torch.cuda.synchronize(device)
            acc["dt"] += time.perf_counter() - t0
            r = bpmod.last_run
            acc["dev_ms"] += r.device_ms
            acc["updates"] += r.updates
            acc["h2d"] += r.h2d_bytes
            acc["d2h"] += r.d2h_bytes
            acc["n_out"] += len(out) if not isinstance(out, list) else sum(len(o) for o in out)
            acc["slices"] += md.num_backpropagated_slices
            for k, v in (r.profile or {}).items():
                a = acc["prof"].setdefault(k, {"launches": 0, "passes": 0, "ms": 0.0, "term_updates": 0})
                for f in a:
                    a[f] += v[f]
        return acc

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()  # before the warm-up: by the timed region it is in its steady polling
        sampler.wait_ready()
    for _ in range(args.warmup):
        one_step()
    barrier()
    # a generation-2 garbage collection over the circuit objects costs ~35 ms, more than a whole step:
    # collect now and keep the collector off while timing (what `timeit` does)
    import gc

    gc.collect()
    gc.disable()
    sampler.mark_begin()
    e2e_s, dev_ms, updates = 0.0, 0.0, 0
    h2d = d2h = 0
    prof_tot: dict = {}
    n_out = n_slices = 0
    wall0 = time.perf_counter()
    poll_every = max(1, int(os.environ.get("OBP_BENCH_POLL_EVERY", "1")))  # sample the clocks after every n-th step
    for _ in range(args.steps):
        acc = one_step()
        if rank == 0 and (_ % poll_every) == poll_every - 1:
            sampler.poll_once()  # NVML: between steps, never inside one (see ClockSampler)
        if os.environ.get("OBP_BENCH_VERBOSE"):
            print(f"[bench] rank {rank} step e2e {acc['dt'] * 1e3:.2f} ms device {acc['dev_ms']:.2f} ms", file=sys.stderr, flush=True)
        e2e_s += acc["dt"]
        dev_ms += acc["dev_ms"]
        updates += acc["updates"]
        h2d += acc["h2d"]
        d2h += acc["d2h"]
        for k, v in acc["prof"].items():
            a = prof_tot.setdefault(k, {"launches": 0, "passes": 0, "ms": 0.0, "term_updates": 0})
            for f in a:
                a[f] += v[f]
        n_out, n_slices = acc["n_out"], acc["slices"]
    barrier()
    wall = time.perf_counter() - wall0
    gc.enable()
    clocks = sampler.stop() if rank == 0 else None

    # max over ranks of the times, sum over ranks of the work
    stats = torch.tensor([e2e_s, dev_ms, float(updates), float(h2d), float(d2h)], dtype=torch.float64, device=device)
    if world > 1:
        mx = stats.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stats.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        e2e_s, dev_ms = float(mx[0]), float(mx[1])
        # a sharded run already reports the whole job's update count on every rank
        updates_all = float(mx[2]) if sharded else float(sm[2])
        h2d_all, d2h_all = float(sm[3]), float(sm[4])
    else:
        updates_all, h2d_all, d2h_all = float(updates), float(h2d), float(d2h)

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
                peaks = json.load(fh)
        except OSError:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        btg = algorithmic_bytes_per_update(wls[0].num_qubits)
        rot = prof_tot.get("rotate", {"launches": 0, "passes": 0, "ms": 0.0, "term_updates": 0})
        achieved = (rot["term_updates"] * btg / (rot["ms"] * 1e-3) / 1e9) if rot["ms"] > 0 else None
        kernel_ms = {k: round(v["ms"] / max(1, args.steps), 3) for k, v in prof_tot.items()}
        launches = int(sum(v["launches"] for v in prof_tot.values()) / max(1, args.steps))
        line = {
            "metric": METRIC,
            "value": updates_all / (dev_ms * 1e-3),
            "unit": UNIT,
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": dev_ms / max(1, args.steps),
            "higher_is_better": True,
            "scaling": "strong" if sharded else "weak",
            "vs_baseline": None,
            "dtype": "u64 keys + f64 complex coefficients",
            "data": "synthetic",
            "config": {
                "workload": workload_name(wls),
                "theta": args.theta,
                "max_paulis": args.max_paulis,
                "slices_absorbed": n_slices,
                "terms_out": n_out,
                "updates_per_step": updates_all / max(1, args.steps),
                "parallelism": (
                    f"one observable hash-sharded over {world} GPUs (pairwise P2P record exchange per branching gate)"
                    if sharded
                    else (f"{world} independent observables, one per GPU" if world > 1 else "1 GPU")
                ),
                "l2": "256 MiB buffer rewritten before every step (L2 flush)",
            },
            "e2e": None if args.no_download else {
                "value": updates_all / e2e_s,
                "unit": UNIT,
                "ms_per_step": 1e3 * e2e_s / max(1, args.steps),
                "h2d_bytes_per_step": h2d_all / max(1, args.steps),
                "d2h_bytes_per_step": d2h_all / max(1, args.steps),
            },
            "gpu_launches": launches,
            "kernel_ms_per_step": kernel_ms,
            "roofline": {
                "kernel": "k_rotate",
                "bound": "hbm",
                "achieved": achieved,
                "peak": peak,
                "peak_source": peak_src,
                "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "frac_of_nominal_8000": (achieved / 8000.0) if achieved else None,
                "bytes_per_term_gate_update": btg,
                "passes": rot["passes"],
                "avg_pass_ms": rot["ms"] / max(1, rot["passes"]),
                "timing": "tables <= 128 qubits: %globaltimer stamps at the grid barriers of the persistent program "
                "kernel; wider tables: CUDA events around the k_rotate launches",
                # one `ncu --set full` capture of the persistent program kernel on this workload (theta_h = 0.3,
                # the 88-rotation slice entered with 94 842 terms: 23 070 231 term-gate updates in one launch)
                "traffic": {
                    "dram_bytes_per_launch": 171.7e6,
                    "algorithmic_bytes_per_launch": 23070231 * 96.0,
                    "kernel": "k_program<2>",
                    "note": "the table (<= 1.5e6 rows x 56 B) stays in the 126 MB L2: DRAM moves 8 % of the algorithmic bytes; "
                            "the pass is bound by the dependent L2 round trips between grid barriers, not by bytes",
                    "source": "profiles/r1i_k_program_ncu.txt (dram__bytes_read.sum + dram__bytes_write.sum)",
                } if args.workload == "config2" and btg == 96 else None,
            },
            "clocks": clocks,
            "wall_s": wall,
        }
        if world == 1 and not args.no_large_table and args.workload == "config2":
            line["roofline_large_table"] = large_table_roofline(backpropagate, bpmod, peak)
        if world == 1 and not args.no_cpu_baseline:
            u, t = cpu_sample(wls, args.cpu_updates)
            line["cpu_baseline"] = {
                "value": u / t,
                "unit": UNIT,
                "cores": 1,
                "kind": "port",
                "sample": f"first {u} term-gate updates of the same workload ({t:.1f} s of the numpy oracle port, 1 thread)",
            }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["engine", "reference"], default="engine")
    ap.add_argument("--workload", default="config2")
    ap.add_argument("--theta", default=f"0.1,0.3,{np.pi / 4:.16f}", help="rotation angle(s), comma separated (config2: theta_h sweep)")
    ap.add_argument("--no-profile", action="store_true", help="do not bracket kernel launches with CUDA events")
    ap.add_argument("--max-paulis", type=int, default=1_000_000)
    ap.add_argument("--cpu-updates", type=int, default=3_000_000, help="bound of the CPU sample (term-gate updates)")
    ap.add_argument("--depth", type=int, default=0, help="circuit depth of config4/config5 (0 = 200 / 40)")
    ap.add_argument("--no-download", action="store_true", help="leave the result on the device(s): for 1e8-term tables whose "
                    "host copy would be tens of GB (e2e is then not an end-to-end number and is reported as null)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-large-table", action="store_true", help="skip the extra HBM-regime roofline pass")
    ap.add_argument("--sharded", action="store_true", help="N>1: shard ONE observable over all ranks instead of one observable per rank")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_engine(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
