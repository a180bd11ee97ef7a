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
observables partition the path with no data-path collective): weak scaling — that is the
top-level line.  The same run then measures the north-star split as the ``sharded`` object: BASELINE
config 4 (1000-qubit near-Clifford brickwork, ``max_paulis=1e8``) with ONE observable hash-sharded over
the N ranks (pairwise P2P record exchange per branching gate), after a parity check of the sharded path
against the oracle over NCCL; ``ms_1gpu`` is the same job on one unsharded table (rank 0 alone, same
run), ``speedup`` their ratio.  At N = 1 the object holds the single-table time only.  ``--sharded``
instead makes the top-level line itself a sharded run of ``--workload``.

N = 1 also reports ``other_configs``: BASELINE configs 1, 3, 4 (2e7 terms) and 5 (1e6) through the same
public call, outside the headline's timed region.
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


def bench_config(args, wls) -> dict:
    """What is measured — the same dictionary in the engine arm and in the reference arm."""
    return {"workload": workload_name(wls), "theta": args.theta, "max_paulis": args.max_paulis}


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
    full = None
    if args.workload == "config2" and not args.no_full_run:
        # one member of the sweep run to COMPLETION (same work as the engine's `other_configs["config2@theta=pi/4"]`):
        # a wall-time comparison beside the rate comparison of the bounded samples
        from qiskit_addon_obp_b200 import workloads as wl_mod

        w = wl_mod.config2_kicked_ising(theta_h=np.pi / 4, steps=20, max_paulis=args.max_paulis)
        t0 = time.perf_counter()
        u, _ = cpu_sample([w], 10**15)
        t_full = time.perf_counter() - t0
        full = {"workload": w.name, "updates": u, "wall_s": t_full, "updates_per_s": u / t_full, "same_work": True,
                "engine_counterpart": "other_configs['config2@theta=pi/4'] of the engine arm's line (e2e_ms)"}
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
        "config": bench_config(args, wls),
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
        "full_workload": full,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------
# engine arm
# ---------------------------------------------------------------------------------------------
def flush_l2(torch, device, buf):
    buf.add_(1)  # writes 256 MiB > 126 MB L2
    torch.cuda.synchronize(device)


def rotation_roofline(prof: dict, btg: int, peak: float) -> dict:
    """Roofline entry of the dominant rotation kernel of a run.

    Two kernels apply rotations: ``k_rotate`` (one pass over the table per gate; per-class time from %globaltimer
    stamps / CUDA events) and the slice-in-shared-memory run ``k_seg_main`` (+ its partition kernels; CUDA events
    around the launches of a run).  ``achieved`` is SURVEY 8(d)'s algorithmic figure for the dominant one: B_tg bytes
    per term-gate update x the updates its launches processed / their device time.  For ``k_seg_main`` that is the
    traffic a gate-by-gate kernel would need for the same work — the run itself reads and writes the table ONCE
    per launch (all gates of the slice are applied in shared memory), so the figure can exceed the HBM peak and is
    an equivalent rate, not DRAM bytes; ``gates_per_launch`` says by how much the real traffic is smaller."""
    zero = {"launches": 0, "passes": 0, "ms": 0.0, "term_updates": 0}
    rot, seg = prof.get("rotate", zero), prof.get("segment", zero)
    top, name = (seg, "k_seg_main") if seg["ms"] >= rot["ms"] else (rot, "k_rotate")
    achieved = (top["term_updates"] * btg / (top["ms"] * 1e-3) / 1e9) if top["ms"] > 0 else None
    both_ms = rot["ms"] + seg["ms"]
    out = {
        "kernel": name,
        "bound": "hbm",
        "achieved": achieved,
        "peak": peak,
        "unit": "GB/s",
        "frac": (achieved / peak) if achieved else None,
        "bytes_per_term_gate_update": btg,
        "passes": top["passes"],
        "avg_pass_ms": top["ms"] / max(1, top["passes"]),
        "term_updates": top["term_updates"],
        "ms": top["ms"],
        "all_rotations": {
            "k_rotate": {"ms": rot["ms"], "passes": rot["passes"], "term_updates": rot["term_updates"]},
            "k_seg_main": {"ms": seg["ms"], "gates": seg["passes"], "launches": seg["launches"],
                           "term_updates": seg["term_updates"]},
            "achieved_gbs_both": ((rot["term_updates"] + seg["term_updates"]) * btg / (both_ms * 1e-3) / 1e9) if both_ms > 0 else None,
        },
        "traffic": None,
    }
    if name == "k_seg_main":
        runs = max(1, seg["launches"] // 4)
        out["gates_per_launch"] = seg["passes"] / runs
        out["note"] = ("slice-in-shared-memory run: the table is read and written once per launch, the gates of the slice are "
                       "applied in shared memory; achieved is the equivalent gate-by-gate traffic (B_tg per term-gate update), "
                       "not DRAM bytes")
    return out


def large_table_roofline(backpropagate, bpmod, peak: float) -> dict:
    """The same k_rotate kernel where it IS bound by HBM: config 2 (theta_h = 0.3) allowed to grow to
    2e7 terms (peak ~2.7e7 rows x 56 B + index: 20x the L2), one untimed-for-`value` pass, per-class device
    time from CUDA events around the k_rotate launches."""
    from qiskit_addon_obp_b200 import workloads as wl

    w = wl.config2_kicked_ising(theta_h=0.3, steps=20, max_paulis=20_000_000)
    backpropagate(w.observables, w.slices, **w.kwargs())  # warm-up: table allocation, slice programs
    backpropagate(w.observables, w.slices, **w.kwargs())
    btg = algorithmic_bytes_per_update(w.num_qubits)
    out = rotation_roofline(bpmod.last_run.profile or {}, btg, peak)
    out["workload"] = w.name
    out["updates_per_s_whole_run"] = bpmod.last_run.updates / (bpmod.last_run.device_ms * 1e-3)
    out["segments"] = dict(zip(("runs", "runs_cut_into_pieces", "fallbacks_to_per_gate_kernels"), bpmod.last_run.segments))
    out["traffic_source"] = "profiles/README.md (ncu --set full captures of the rotation kernels on tables of this size)"
    return out


def other_configs(backpropagate, bpmod, torch, device, flush_buf, peak: float) -> dict:
    """BASELINE configs 1, 3, 4 and 5 through the same public call (N = 1, outside the headline's timed
    region): one warm-up call, then `reps` timed calls each; device-resident and end-to-end time, updates/s
    and the algorithmic HBM fraction of the kernel class that took the most device time.  Config 1 is small
    enough for the CPU port to run the WHOLE job: its wall time is measured beside the engine's (same work)."""
    from qiskit_addon_obp_b200 import workloads as wl

    cases = [
        ("config1", wl.config1_xy_chain(), 5, True),
        ("config3", wl.config3_square_lattice(), 2, True),
        ("config4@2e7", wl.config4_near_clifford(depth=200, max_paulis=20_000_000), 2, False),
        ("config5@1e6", wl.config5_magic_sweep(theta=np.pi / 16, depth=40, max_paulis=1_000_000), 2, True),
        # the one member of the headline sweep the CPU arm runs to completion (`full_workload` of --impl reference)
        ("config2@theta=pi/4", wl.config2_kicked_ising(theta_h=np.pi / 4, steps=20, max_paulis=1_000_000), 3, True),
    ]
    out = {}
    for name, w, reps, download in cases:
        os.environ["OBP_B200_DOWNLOAD"] = "1" if download else "0"
        try:
            backpropagate(w.observables, w.slices, **w.kwargs())  # warm-up: tables, compiled slices
            dev_ms = wall = 0.0
            updates = 0
            prof: dict = {}
            res = None
            for _ in range(reps):
                res = None  # (drop the previous result first: its page-locked buffers go back to the host allocator's cache)
                flush_l2(torch, device, flush_buf)
                t0 = time.perf_counter()
                res, rest, md = backpropagate(w.observables, w.slices, **w.kwargs())
                torch.cuda.synchronize(device)
                wall += time.perf_counter() - t0
                r = bpmod.last_run
                dev_ms += r.device_ms
                updates += r.updates
                for k, v in (r.profile or {}).items():
                    a = prof.setdefault(k, {"ms": 0.0, "term_updates": 0})
                    a["ms"] += v["ms"]
                    a["term_updates"] += v["term_updates"]
            btg = algorithmic_bytes_per_update(w.num_qubits)
            top = max((k for k in prof if k != "program"), key=lambda k: prof[k]["ms"], default=None)
            entry = {
                "workload": w.name,
                "slices_absorbed": md.num_backpropagated_slices,
                "terms_out": len(res) if download else int(md.backpropagation_history[md.num_backpropagated_slices - 1].num_paulis[0]),
                "updates_per_call": updates / reps,
                "device_ms": dev_ms / reps,
                "e2e_ms": 1e3 * wall / reps if download else None,
                "updates_per_s": updates / (dev_ms * 1e-3) if dev_ms > 0 else None,
                "kernel_ms": {k: round(v["ms"] / reps, 3) for k, v in prof.items() if v["ms"] > 0},
            }
            if top is not None and prof[top]["ms"] > 0 and prof[top]["term_updates"]:
                dom = {"class": top, "ms": round(prof[top]["ms"] / reps, 3),
                       "term_gate_updates_per_s": prof[top]["term_updates"] / (prof[top]["ms"] * 1e-3)}
                if top in ("rotate", "segment"):  # B_tg bytes per update: the algorithmic (gate-by-gate) traffic
                    ach = prof[top]["term_updates"] * btg / (prof[top]["ms"] * 1e-3) / 1e9
                    dom.update({"achieved_gbs_algorithmic": ach, "frac": ach / peak, "bytes_per_term_gate_update": btg})
                else:  # a fused Clifford pass applies a whole batch of gates per sweep over the table: updates/s is
                    dom["note"] = "no per-update roofline: one pass over the table serves every gate fused into it"  # not bytes/s
                entry["dominant_kernel"] = dom
            if name == "config1":
                t0 = time.perf_counter()
                u, _ = cpu_sample([w], 10**12)
                t_cpu = time.perf_counter() - t0
                entry["cpu_full_run"] = {"kind": "port", "cores": 1, "wall_ms": 1e3 * t_cpu, "updates": u,
                                         "same_work": True, "e2e_speedup": t_cpu / (wall / reps)}
            out[name] = entry
        except Exception as exc:  # noqa: BLE001 - a side measurement must not take the headline line down
            out[name] = {"workload": w.name, "error": f"{type(exc).__name__}: {exc}"}
    os.environ.pop("OBP_B200_DOWNLOAD", None)
    return out


def sharded_parity_check(backpropagate, bpmod) -> tuple[bool, str]:
    """The sharded path against the oracle over NCCL, on every rank (the oracle only checks here): small circuits
    sharded from the first commit, and one that starts replicated and is split mid-run."""
    from oracle import obp_oracle as oracle
    from qiskit_addon_obp_b200 import workloads as wl
    from qiskit_addon_obp_b200.ir import QuantumCircuit, SparsePauliOp

    def window(nq, rounds):
        lo = max(0, min(nq, 70) - 12)
        win = list(range(lo, min(nq, lo + 12)))
        slices = []
        for k in range(rounds):
            qc = QuantumCircuit(nq)
            for q in win:
                qc.rx(0.5 + 0.01 * q, q)
            slices.append(qc)
            qc = QuantumCircuit(nq)
            for a, b in zip(win[k % 2 :: 2], win[k % 2 + 1 :: 2]):
                qc.rzz(0.8, a, b)
                qc.cx(a, b)
            slices.append(qc)
            qc = QuantumCircuit(nq)
            for q in win:
                qc.h(q)
                qc.ry(0.4, q)
            slices.append(qc)
        return SparsePauliOp.from_sparse_list([("Z", [win[5]], 1.0), ("XY", [win[2], win[3]], 0.7)], nq), slices

    def same(got, want, md, md_w) -> str:
        a, b = got.as_dict(), want.as_dict()
        if set(a) != set(b):
            return f"key sets differ ({len(a)} vs {len(b)})"
        err = max((abs(a[k] - b[k]) - 1e-10 * abs(b[k]) for k in a), default=0.0)
        if err > 1e-12:
            return f"coefficients differ by {err:.3e}"
        if md.num_backpropagated_slices != md_w.num_backpropagated_slices:
            return "slices absorbed differ"
        for g, w in zip(md.backpropagation_history, md_w.backpropagation_history):
            if g.num_paulis != w.num_paulis or g.sum_paulis != w.sum_paulis:
                return "per-slice term counts differ"
            if max(abs(x - y) for x, y in zip(g.slice_errors, w.slice_errors)) > 1e-9:
                return "slice errors differ"
        return ""

    cases = []
    c1 = wl.config1_xy_chain()
    cases.append(("config1", "0", c1.observables, c1.slices, c1.kwargs()))
    c4 = wl.config4_near_clifford(num_qubits=1000, depth=10, t_prob=0.01, max_paulis=4000, per_slice=1e-4)
    cases.append(("config4 small", "0", c4.observables, c4.slices, c4.kwargs()))
    obs, sl = window(70, 4)
    cases.append(("70q window, sharded at once", "0", obs, sl, {}))
    cases.append(("70q window, split mid-run", "300", obs, sl, {}))
    cases.append(("70q window, split mid-run, hash shard function", "300/noplan", obs, sl, {}))
    exchanges = 0
    saved_download = os.environ.get("OBP_B200_DOWNLOAD")
    os.environ["OBP_B200_DOWNLOAD"] = "1"  # the check needs the operators back
    try:
        for name, min_terms, o, sl, kw in cases:
            os.environ["OBP_B200_SHARD_PLAN"] = "0" if min_terms.endswith("/noplan") else "1"
            min_terms = min_terms.split("/")[0]
            os.environ["OBP_B200_SHARD_MIN_TERMS"] = min_terms
            got, rest, md = backpropagate(o, sl, **kw)
            exchanges += bpmod.last_run.exchanges
            want, rest_w, md_w = oracle.backpropagate(o, sl, **kw)
            msg = same(got, want, md, md_w) if len(rest) == len(rest_w) else "remaining slices differ"
            if msg:
                return False, f"{name}: {msg}"
        if exchanges == 0:
            return False, "no record exchange took place"
        return True, f"{len(cases)} cases vs the oracle, {exchanges} record exchanges"
    except Exception as exc:  # noqa: BLE001
        return False, f"{type(exc).__name__}: {exc}"
    finally:
        os.environ.pop("OBP_B200_SHARD_MIN_TERMS", None)
        os.environ.pop("OBP_B200_SHARD_PLAN", None)
        if saved_download is None:
            os.environ.pop("OBP_B200_DOWNLOAD", None)
        else:
            os.environ["OBP_B200_DOWNLOAD"] = saved_download


def sharded_block(args, rank: int, world: int, device, torch, dist, backpropagate, bpmod, flush_buf) -> dict | None:
    """The north-star multi-GPU split, measured in the same run (see the module docstring)."""
    from qiskit_addon_obp_b200 import engine
    from qiskit_addon_obp_b200 import sharded as sh
    from qiskit_addon_obp_b200 import workloads as wl

    w = wl.config4_near_clifford(depth=200, max_paulis=args.sharded_max_paulis)
    os.environ["OBP_B200_DOWNLOAD"] = "0"  # 1e8 terms x 272 B: the result stays on the device(s)
    os.environ.pop("OBP_B200_SHARDED", None)
    engine.clear_pool()

    def barrier():
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    def run(n_warm: int, n_steps: int) -> dict:
        acc = {"dev_ms": 0.0, "wall_ms": 0.0, "updates": 0, "wasted": 0, "prof": {}, "exchanges": 0, "records": 0, "slices": 0, "peak": 0}
        for k in range(n_warm + n_steps):
            flush_l2(torch, device, flush_buf)
            barrier()
            t0 = time.perf_counter()
            _, _, md = backpropagate(w.observables, w.slices, **w.kwargs())
            torch.cuda.synchronize(device)
            barrier()
            if k < n_warm:
                continue
            r = bpmod.last_run
            acc["wall_ms"] += 1e3 * (time.perf_counter() - t0)
            acc["dev_ms"] += r.device_ms
            acc["updates"] += r.updates
            acc["wasted"] += r.wasted_updates
            acc["exchanges"] += r.exchanges
            acc["records"] += r.exchanged_records
            acc["slices"] = md.num_backpropagated_slices
            acc["peak"] = max(s.num_paulis[0] for s in md.backpropagation_history)
            for c, v in (r.profile or {}).items():
                acc["prof"][c] = acc["prof"].get(c, 0.0) + v["ms"]
        for f in ("dev_ms", "wall_ms", "updates", "wasted", "exchanges", "records"):
            acc[f] /= n_steps
        acc["prof"] = {c: round(v / n_steps, 3) for c, v in acc["prof"].items() if v > 0}
        return acc

    out: dict = {"workload": w.name, "n_gpus": world}
    try:
        # ---- one unsharded table: rank 0 alone (its peers wait at the barrier inside run()) ----------------
        single = None
        if rank == 0 or world == 1:
            pass
        # every rank enters run() (it holds the barriers); only rank 0 does the work
        if world == 1:
            single = run(1, 2)
        else:
            single = _rank0_only(run, rank, barrier, 1, 2, w, backpropagate)
        engine.clear_pool()
        if world == 1:
            out.update({"ms_1gpu": single["dev_ms"], "ms_per_step": single["dev_ms"], "speedup": 1.0,
                        "updates_per_step": single["updates"], "slices_absorbed": single["slices"],
                        "peak_terms": single["peak"], "kernel_ms_1gpu": single["prof"],
                        "note": "N = 1: one unsharded table; run with --gpus N for the hash-sharded table"})
            return out
        # ---- the same job, one observable hash-sharded over all ranks ----------------------------------------
        os.environ["OBP_B200_SHARDED"] = "1"
        ok, detail = sharded_parity_check(backpropagate, bpmod)
        flag = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        shard = run(1, 3)
        t = torch.tensor([shard["dev_ms"], shard["wall_ms"]] + [shard["prof"].get(c, 0.0) for c in engine.OBP_K_NAMES],
                         dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms1 = torch.tensor([single["dev_ms"] if single else 0.0, float(single["updates"]) if single else 0.0],
                           dtype=torch.float64, device=device)
        dist.all_reduce(ms1, op=dist.ReduceOp.MAX)
        kernel_ms = {c: round(float(v), 3) for c, v in zip(engine.OBP_K_NAMES, t[2:]) if float(v) > 0}
        dev_ms = float(t[0])
        out.update({
            "ms_1gpu": float(ms1[0]),
            "ms_per_step": dev_ms,
            "wall_ms_per_step": float(t[1]),
            "speedup": float(ms1[0]) / dev_ms if dev_ms > 0 else None,
            "updates_per_step": shard["updates"],
            "wasted_updates_per_step": shard["wasted"],
            "updates_per_s": shard["updates"] / (dev_ms * 1e-3) if dev_ms > 0 else None,
            "slices_absorbed": shard["slices"],
            "peak_terms": shard["peak"],
            "parity_ok": bool(float(flag[0]) > 0),
            "parity_detail": detail,
            "phases_ms": {
                "clifford": kernel_ms.get("clifford", 0.0),
                "rotate_emit_merge_finish": kernel_ms.get("rotate", 0.0),
                "other_kernels": round(sum(v for c, v in kernel_ms.items() if c not in ("clifford", "rotate")), 3),
                "host_and_wait": round(max(0.0, dev_ms - sum(kernel_ms.values())), 3),
                "note": "max over ranks of CUDA-event time per kernel class; a split rotation is three kernels (emit into the "
                        "peer's inbox over NVLink, merge, finish) timed as one class; host_and_wait = device-timed span minus the kernel classes",
            },
            "kernel_ms_1gpu": single["prof"] if single else None,
            "exchanges_per_step": shard["exchanges"],
            "exchanged_records_per_step": shard["records"],
            "shard_min_terms": sh.ShardedTable.split_min_env(),
            "shard_function": "planned at the split: owner bits of every upcoming rotation generator are zero, so a rotation "
                              "never moves a term to another rank (exchanges_per_step counts the ones outside the plan)",
            "timing": "CUDA events on each rank's stream from the observable upload to the end of the last slice, max over ranks; "
                      "result left on the devices (OBP_B200_DOWNLOAD=0)",
        })
        sh.close_all()
        return out
    except Exception as exc:  # noqa: BLE001 - report, never take the headline line down
        out["error"] = f"{type(exc).__name__}: {exc}"
        return out
    finally:
        os.environ.pop("OBP_B200_DOWNLOAD", None)
        os.environ.pop("OBP_B200_SHARDED", None)


def _rank0_only(run, rank, barrier, n_warm, n_steps, w, backpropagate):
    """run() on rank 0 while the other ranks only keep its barriers company."""
    if rank == 0:
        return run(n_warm, n_steps)
    for _ in range(2 * (n_warm + n_steps)):
        barrier()
    return None


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
        acc = {"dt": 0.0, "dev_ms": 0.0, "updates": 0, "wasted": 0, "h2d": 0, "d2h": 0, "n_out": 0, "slices": 0, "prof": {}, "seg": (0, 0, 0)}
        for wl in wls:
            t0 = time.perf_counter()
            out, rest, md = backpropagate(wl.observables, wl.slices, **wl.kwargs())
            torch.cuda.synchronize(device)
            acc["dt"] += time.perf_counter() - t0
            r = bpmod.last_run
            acc["dev_ms"] += r.device_ms
            acc["updates"] += r.updates
            acc["wasted"] += getattr(r, "wasted_updates", 0)
            acc["h2d"] += r.h2d_bytes
            acc["d2h"] += r.d2h_bytes
            acc["n_out"] += len(out) if not isinstance(out, list) else sum(len(o) for o in out)
            acc["slices"] += md.num_backpropagated_slices
            acc["seg"] = tuple(a + b for a, b in zip(acc["seg"], getattr(r, "segments", (0, 0, 0))))
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
    h2d = d2h = wasted = 0
    prof_tot: dict = {}
    seg_tot = (0, 0, 0)
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
        wasted += acc["wasted"]
        h2d += acc["h2d"]
        d2h += acc["d2h"]
        for k, v in acc["prof"].items():
            a = prof_tot.setdefault(k, {"launches": 0, "passes": 0, "ms": 0.0, "term_updates": 0})
            for f in a:
                a[f] += v[f]
        n_out, n_slices = acc["n_out"], acc["slices"]
        seg_tot = tuple(a + b for a, b in zip(seg_tot, acc["seg"]))
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
        roof = rotation_roofline(prof_tot, btg, peak)
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
            "config": bench_config(args, wls),
            "run": {
                "slices_absorbed": n_slices,
                "terms_out": n_out,
                "updates_per_step": updates_all / max(1, args.steps),
                "wasted_updates_per_step": wasted / max(1, args.steps),
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
            "roofline": dict(roof, peak_source=peak_src, frac_of_nominal_8000=(roof["achieved"] / 8000.0) if roof["achieved"] else None,
                             timing="k_rotate inside the persistent program kernel: %globaltimer stamps at its grid barriers; stand-alone "
                                    "k_rotate and the k_seg_* launches of a run: CUDA events on the table's stream",
                             traffic_source="profiles/README.md (ncu --set full captures of the rotation kernels, per launch)"),
            "segments": dict(zip(("runs", "runs_cut_into_pieces", "fallbacks_to_per_gate_kernels"), seg_tot)),
            "clocks": clocks,
            "wall_s": wall,
        }
        if world == 1 and not args.no_large_table and args.workload == "config2":
            line["roofline_large_table"] = large_table_roofline(backpropagate, bpmod, peak)
        if world == 1 and not args.no_other_configs and args.workload == "config2":
            line["other_configs"] = other_configs(backpropagate, bpmod, torch, device, flush_buf, peak)
        if world == 1 and not args.no_cpu_baseline:
            u, t = cpu_sample(wls, args.cpu_updates)
            line["cpu_baseline"] = {
                "value": u / t,
                "unit": UNIT,
                "cores": 1,
                "kind": "port",
                "sample": f"first {u} term-gate updates of the same workload ({t:.1f} s of the numpy oracle port, 1 thread)",
            }
    else:
        line = None
    if not args.no_sharded_block and not sharded and args.workload == "config2":
        blk = sharded_block(args, rank, world, device, torch, dist if world > 1 else None, backpropagate, bpmod, flush_buf)
        if rank == 0:
            line["sharded"] = blk
    if rank == 0:
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
    ap.add_argument("--no-full-run", action="store_true", help="--impl reference: skip the one workload run to completion (~70 s)")
    ap.add_argument("--no-large-table", action="store_true", help="skip the extra HBM-regime roofline pass")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the configs 1/3/4/5 side measurements (N = 1)")
    ap.add_argument("--no-sharded-block", action="store_true", help="skip the config-4 hash-sharded strong-scaling measurement")
    ap.add_argument("--sharded-max-paulis", type=int, default=100_000_000, help="max_paulis of the sharded config-4 measurement")
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
