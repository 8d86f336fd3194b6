#!/usr/bin/env python
"""bench.py -- BN254 Groth16 proofs verified per second (BASELINE.json metric) on N B200s.

A step = one pass of the hot path over one batch: `--batch` (default 2^16) trapdoor-simulated Groth16 proofs
per GPU, 2 public inputs, 50 % corrupted (BASELINE.json configs[1]).  The batch shards by proof index: one
process per GPU, no collective on the data path; the only exchange is the final gather of verdict bits.

  value   proofs/s with the batch already resident in HBM (kernel time, CUDA events on the launching stream)
  e2e     proofs/s through the public API (Groth16Verifier.verify_batch -> bn254v_groth16_verify_batch) from
          pinned HOST buffers: H2D of proofs+inputs and D2H of the status bytes inside the timed region
  roofline  int32 multiply-add pipe: algorithmic limb-MACs per proof x proofs / kernel time, against
          the IMAD.WIDE issue rate measured live by bn254v_imad_peak (MEASURED_PEAKS.json has no integer peak)
  cpu_baseline / --impl reference   the oracle's C++ restatement of the reference algorithm (the Rust crate
          cannot be built here: no Rust toolchain, bn dependency not on disk) on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "groth16_proofs_verified_per_sec"
UNIT = "proofs/s"
SEED = 20240607
# Algorithmic work per proof, counted by the instrumented oracle executing the GPU-shaped algorithm
# (tests/test_workcount.py, DESIGN.md "work per unit"): Fp multiplications x 136 limb-MACs (8x32-bit CIOS).
MACS_PER_FPMUL = 136


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1 << 16, help="proofs per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="proofs in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the PlonK / raw pairing-product side measurements")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) < 8:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def ncu_traffic(n):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        t = json.load(open(p))["k_groth16_miller"]
        return t["dram_bytes_per_launch"] if t["batch"] == n else None
    except Exception:
        return None


def work_per_proof():
    p = os.path.join(ROOT, "profiles", "workcount.json")
    if os.path.exists(p):
        return json.load(open(p))
    return {"groth16_fp_mul": 32500, "source": "SURVEY.md Appendix D estimate (workcount.json missing)"}


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's C++ restatement on the host cores
# --------------------------------------------------------------------------------------------------
def cpu_reference_run(vk, proofs, inputs, threads):
    """Times oracle/ (test infrastructure, the CHECKER) verifying `proofs` on `threads` host threads.
    Returns (seconds, status)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_cpu  # oracle/ref_cpu.py: ctypes loader of oracle/_build/libbn254ref.so
    return ref_cpu.groth16_verify_batch(vk, proofs, inputs, threads)


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    """`--impl reference`: the reference algorithm (C++ restatement, reference's shape: VK parsed per call,
    4 Miller loops + 2 final exponentiations, naive MSM) on all host cores, same workload and metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_cpu
    cores = host_cores()
    sample = args.cpu_sample or max(cores * 256, 512)
    vk, proofs, inputs, expected = ref_cpu.groth16_synth(SEED, sample)
    times = []
    for it in range(args.warmup + args.steps):
        dt, status = ref_cpu.groth16_verify_batch(vk, proofs, inputs, cores)
        assert (status == expected).all()
        if it >= args.warmup:
            times.append(dt)
    total = sum(times)
    value = sample * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": "groth16 trapdoor-simulated proofs, 2 public inputs, 50%% corrupted "
                               "(BASELINE.json configs[1]); bounded sample of %d proofs per step" % sample,
                   "seed": SEED},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d proofs per step, %d steps" % (sample, args.steps),
                         "note": "C++ restatement of the reference algorithm in the reference's shape; the Rust "
                                 "crate cannot be built here (no Rust toolchain, bn dependency not on disk)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def secondary_workloads(pkg, work):
    """Side measurements on rank 0 (end to end through the C ABI from host buffers, wall clock around the synchronous
    call): BASELINE.json configs[2] (2^14 PlonK proofs from the bundled fixtures, 50 % mutated) and configs[3] scaled
    to 2^17 raw 4-pair products.  Reported next to the headline; the headline metric stays Groth16 configs[1]."""
    import numpy as np
    import workloads
    out = {}
    n = 1 << 14
    proofs, inputs, rnd, expected = workloads.plonk_workload(n, seed=3)
    vk = workloads.plonk_vk_bytes()
    pkg.PlonkVerifier.verify_batch(proofs[:512], vk, inputs[:512], rnd=rnd[:512])  # VK load + warm-up
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        st = pkg.PlonkVerifier.verify_batch(proofs, vk, inputs, rnd=rnd)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    assert (st == expected).all(), "PlonK status mismatch"
    out["plonk"] = {"workload": "2^14 PlonK proofs = 4 bundled SP1 fixtures replicated, 50% mutated (BASELINE configs[2])",
                    "e2e_proofs_per_sec": n / best, "ms": best * 1e3,
                    "field_mults_full_path": work.get("plonk_full_path_mul"),
                    "note": "mutated proofs split between early reject (OpeningPolyMismatch) and full-path reject"}
    p2, i2, r2, e2 = workloads.plonk_workload(n, seed=4, late_reject_only=True)
    t0 = time.perf_counter()
    st = pkg.PlonkVerifier.verify_batch(p2, vk, i2, rnd=r2)
    dt = time.perf_counter() - t0
    assert (st == e2).all()
    out["plonk"]["late_reject_only_proofs_per_sec"] = n / dt
    # the same records 8x: 2^17 proofs fill the GPU (2^14 leaves the final pairing stage at <1 warp per SMSP)
    big = 8
    pb, ib, rb, eb = (np.tile(x, (big,) + (1,) * (x.ndim - 1)) for x in (proofs, inputs, rnd, expected))
    dt = None
    for _ in range(2):
        t0 = time.perf_counter()
        st = pkg.PlonkVerifier.verify_batch(pb, vk, ib, rnd=rb)
        d1 = time.perf_counter() - t0
        dt = d1 if dt is None else min(dt, d1)
    assert (st == eb).all()
    out["plonk"]["e2e_proofs_per_sec_2e17_batch"] = n * big / dt
    try:  # CPU side by side: the oracle's C++ restatement of PlonkVerifier::verify on the same records
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_cpu
        cores = host_cores()
        sample = min(n, max(64 * cores, 256))
        dtc, stc = ref_cpu.plonk_verify_batch(vk, proofs[:sample], inputs[:sample], rnd[:sample], threads=cores)
        assert (stc == expected[:sample]).all(), "CPU PlonK oracle disagrees"
        out["plonk"]["cpu_baseline"] = {"value": sample / dtc, "unit": "proofs/s", "cores": cores, "kind": "port",
                                        "sample": "first %d records of the same batch, one pass, %.1f s" % (sample, dtc)}
    except Exception as e:
        out["plonk"]["cpu_baseline"] = {"value": None, "sample": "unavailable: %r" % (e,)}
    m = 1 << 17
    g1, g2, exp1 = pkg.pairing_synth(11, m, k=4)
    dt = None  # best of 3: the first full-size call also pays the kernel's one-time module and local-memory set-up
    for _ in range(3):
        t0 = time.perf_counter()
        one = pkg.pairing_product_batch(g1, g2, 4)
        d1 = time.perf_counter() - t0
        dt = d1 if dt is None else min(dt, d1)
    assert (one == exp1).all()
    out["pairing_product_k4"] = {"workload": "2^17 random 4-pair sets, all G2 variable (BASELINE configs[3] scaled)",
                                 "e2e_sets_per_sec": m / dt, "pair_miller_loops_per_sec": 4 * m / dt, "ms": dt * 1e3}
    return out


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import numpy as np
    import torch
    import __graft_entry__ as ge

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    pkg = ge.load_package()
    from importlib import import_module
    sharding = import_module("snark_bn254_verifier_b200.sharding")
    pkg.init([local_rank])
    n = args.batch

    # synthetic workload, generated on the device by the library (same definition as the oracle's generator)
    vk, proofs, inputs, expected = pkg.groth16_synth(SEED, n, first_index=rank * n)
    # pinned host staging for the e2e path
    t_proofs = torch.from_numpy(proofs).pin_memory()
    t_inputs = torch.from_numpy(inputs).pin_memory()
    t_status = torch.empty(n, dtype=torch.uint8).pin_memory()
    np_proofs, np_inputs = t_proofs.numpy(), t_inputs.numpy()
    h2d = np_proofs.nbytes + np_inputs.nbytes
    d2h = n

    batch = pkg.Groth16DeviceBatch(vk, proofs, inputs)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak = pkg.imad_peak(2048) if rank == 0 else None

    # ---- kernel-only: inputs resident in HBM ------------------------------------------------------
    for _ in range(args.warmup):
        batch.verify(want_status=False)
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = pkg.launch_count()
    kernel_ms, split_ms = [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flush.zero_()
        torch.cuda.synchronize()
        st, ms = batch.verify(want_status=False)
        kernel_ms.append(ms)
        split_ms.append(pkg.last_kernel_split())
    barrier()
    wall_kernel = time.perf_counter() - t0
    launches = pkg.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    status, _ = batch.verify(want_status=True)
    assert (status == expected).all(), "verdict mismatch against the generator's expected verdicts"
    dev_ms = max_over_ranks(sum(kernel_ms))  # device time of K steps, max over ranks
    value = world * n * args.steps / (dev_ms * 1e-3)

    # ---- end to end through the public API, host buffers ------------------------------------------
    for _ in range(max(1, args.warmup // 2)):
        pkg.Groth16Verifier.verify_batch(np_proofs, vk, np_inputs)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        st = pkg.Groth16Verifier.verify_batch(np_proofs, vk, np_inputs)  # H2D + kernel + D2H, synchronous
        if dist is not None:  # final gather of verdict bits (n/8 bytes per rank) -- the path's only exchange
            verdicts = sharding.gather_verdicts(st, world * n, dist, device="cuda")
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    assert (st == expected).all()
    e2e_value = world * n * args.steps / e2e_s

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    work = work_per_proof()
    macs_per_proof = work.get("groth16_macs", work["groth16_fp_mul"] * MACS_PER_FPMUL)
    step_achieved = macs_per_proof * n * args.steps / (sum(kernel_ms) * 1e-3)  # both launches of this rank's step
    miller_ms = sum(a for a, b in split_ms)
    finish_ms = sum(b for a, b in split_ms)
    two = finish_ms > 0
    # dominant kernel: the Miller-loop launch (its own algorithmic MACs over its own CUDA-event duration)
    dom_macs = work.get("groth16_miller_macs", macs_per_proof) if two else macs_per_proof
    achieved = dom_macs * n * args.steps / (miller_ms * 1e-3)
    roofline = {
        "bound": "int32-imad", "achieved": achieved / 1e12, "peak": peak["wide_mac_per_s"] / 1e12, "unit": "TMAC/s",
        "frac": achieved / peak["wide_mac_per_s"], "traffic": ncu_traffic(n),
        "kernel": "k_groth16_miller<448>" if two else "k_groth16_verify",
        "kernel_ms_per_launch": miller_ms / args.steps, "macs_per_launch": dom_macs * n,
        "share_of_step": miller_ms / sum(kernel_ms),
        "step": {"achieved": step_achieved / 1e12, "frac": step_achieved / peak["wide_mac_per_s"],
                 "kernels": ["k_groth16_miller", "k_groth16_finish"] if two else ["k_groth16_verify"],
                 "finish_ms_per_launch": finish_ms / args.steps,
                 "finish_frac": (work.get("groth16_finish_macs", 0) * n * args.steps / (finish_ms * 1e-3) /
                                 peak["wide_mac_per_s"]) if two else None},
        "macs_per_proof": macs_per_proof, "fp_mul_per_proof": work["groth16_fp_mul"],
        "peak_source": "measured live: bn254v_imad_peak (independent IMAD.WIDE.U32 accumulate chains, 8 warps/SMSP, "
                       "all SMs); MEASURED_PEAKS.json holds no integer peak",
        "peak_imad32_tmacs": peak["lo_mac_per_s"] / 1e12,
        "hbm_gbs_algorithmic": (h2d + d2h) * args.steps / (sum(kernel_ms) * 1e-3) / 1e9,
        "note": "tensor cores unused: carry-chained multiprecision integer arithmetic; HBM traffic negligible",
    }

    cpu = None
    if not args.no_cpu_baseline and world == 1:  # reported at N=1 only
        try:
            cores = host_cores()
            sample = args.cpu_sample or max(cores * 512, 512)
            sample = min(sample, n)
            dt, st_cpu = cpu_reference_run(vk, proofs[:sample], inputs[:sample], cores)
            assert (st_cpu == expected[:sample]).all(), "CPU oracle disagrees with expected verdicts"
            cpu = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "first %d proofs of the same batch, one pass, %.1f s" % (sample, dt)}
        except Exception as e:  # the baseline is a reported number, never a dependency of the product path
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "unavailable: %r" % (e,)}

    extra = {}
    if not args.no_secondary and world == 1:
        extra = secondary_workloads(pkg, work)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": "2^%d trapdoor-simulated Groth16 proofs per GPU, 2 public inputs, 50%% corrupted "
                               "(BASELINE.json configs[1])" % (n.bit_length() - 1),
                   "proofs_per_gpu": n, "seed": SEED, "l2": "flushed between timed iterations (256 MiB write)",
                   "parallelism": "proof-index sharding, %d rank(s), no data-path collective" % world},
        "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "pairings_per_sec": 3 * value,
        "wall_s_kernel_loop": wall_kernel,
        **extra,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
