#!/usr/bin/env python
"""bench.py -- BN254 Groth16 & PlonK proofs verified per second, pairings per second (BASELINE.json metric) on N B200s.

Headline (`value`, `e2e`, `roofline`): BASELINE.json configs[1] -- `--batch` (default 2^16) trapdoor-simulated Groth16
proofs per GPU, 2 public inputs, 50 % corrupted.  The same JSON line carries, at every N, the other configurations of
the metric, each with its own device-timed `value`, `e2e` and `roofline`:

  plonk    configs[2]  2^14 PlonK proofs per GPU from the bundled SP1 fixtures, replicated, 50 % mutated
  pairing  configs[3]  2^20 random 4-pair products per GPU (all G2 variable) -> sets/s and pair-Miller-loops/s
  mixed    configs[4]  2^22 items in all (2^21 Groth16 + 2^21 PlonK, interleaved), sharded by index over the N GPUs
                       (strong scaling), through bn254v_verify_many

A step = one pass of the hot path over one batch.  The batches shard by proof index: one process per GPU, no
collective on the data path; the only exchange is the final gather of the status bytes.

  value     items/s with the batch already resident in HBM (kernel time, CUDA events on the launching stream)
  e2e       items/s through the public API (Groth16Verifier.verify_batch -> bn254v_groth16_verify_batch, ...) from
            pinned HOST buffers: H2D of the records and D2H of the status bytes inside the timed region
  roofline  int32 multiply-add pipe: algorithmic limb-MACs per item (profiles/workcount.json) x items /
            kernel time, against the IMAD.WIDE issue rate measured live by bn254v_imad_peak (MEASURED_PEAKS.json holds
            no integer peak)
  cpu_baseline / --impl reference   the oracle's C++ restatement of the reference algorithm (the Rust crate cannot be
            built here: no Rust toolchain, bn dependency not on disk) on the host cores
  single_process (N > 1)  rank 0 alone drives all N GPUs through one library call (bn254v_init over N devices, the
            in-library sharding loop), N x 2^17 proofs in one step, before the per-rank runs start
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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "groth16_proofs_verified_per_sec"
UNIT = "proofs/s"
SEED = 20240607
MACS_PER_FPMUL = 136  # 8 x 32-bit CIOS Montgomery: 64 product + 72 reduction multiply-adds


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=1 << 16, help="Groth16 proofs per GPU per step")
    ap.add_argument("--plonk-batch", type=int, default=1 << 14, help="PlonK proofs per GPU per step")
    ap.add_argument("--pairing-batch", type=int, default=1 << 20, help="4-pair sets per GPU per step")
    ap.add_argument("--mixed-total", type=int, default=1 << 22, help="items of the mixed workload over ALL GPUs")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workloads", default="groth16,plonk,pairing,mixed,single,allvalid",
                    help="comma list: groth16 (always), plonk, pairing, mixed, single (in-library multi-device, N > 1)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="proofs in the cpu_baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="headline Groth16 workload only")
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


def ncu_traffic(kernel, n):
    """DRAM bytes per launch of a kernel from the committed `ncu --set full` capture (profiles/ncu_traffic.json)."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))[kernel]
        return t["dram_bytes_per_launch"] if t["batch"] == n else None
    except Exception:
        return None


def work_per_proof():
    return json.load(open(os.path.join(ROOT, "profiles", "workcount.json")))


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle's C++ restatement on the host cores
# --------------------------------------------------------------------------------------------------
def _ref_cpu():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import ref_cpu  # oracle/ref_cpu.py: ctypes loader of oracle/_build/libbn254ref.so (the CHECKER, never the product)
    return ref_cpu


def run_reference(args):
    """`--impl reference`: the reference algorithm (C++ restatement, reference's shape: VK parsed per call,
    4 Miller loops + 2 final exponentiations, naive MSM) on all host cores, same workload and metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import workloads
    ref_cpu = _ref_cpu()
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
    extra = {}
    if not args.no_secondary:  # the other configurations of the metric, one bounded pass each
        ps = min(1 << 14, max(64 * cores, 256))
        pp, pi, pr, pe = workloads.plonk_workload(ps, seed=3)
        dt, st = ref_cpu.plonk_verify_batch(workloads.plonk_vk_bytes(), pp, pi, pr, threads=cores)
        assert (st == pe).all()
        extra["plonk"] = {"value": ps / dt, "unit": "proofs/s", "sample": "%d proofs, one pass" % ps}
        import numpy as np
        qs = max(64 * cores, 256)
        g1 = np.tile(proofs[:qs, None, 0:64], (1, 4, 1))
        g2 = np.tile(proofs[:qs, None, 64:192], (1, 4, 1))
        dt, one, _, _ = ref_cpu.pairing_product_batch(g1, g2, 4, threads=cores)
        extra["pairing"] = {"value": qs / dt, "unit": "4-pair sets/s", "pair_miller_loops_per_sec": 4 * qs / dt,
                            "sample": "%d sets, one pass" % qs}
        extra["mixed"] = {"value": 2.0 / (1.0 / value + 1.0 / extra["plonk"]["value"]), "unit": "items/s",
                          "sample": "harmonic mean of the Groth16 and PlonK rates above (1:1 mix)"}
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
        "gpu_launches": 0, **extra,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------------
class Ctx:
    """Per-process state of the B200 arm."""

    def __init__(self, args):
        import torch
        self.args = args
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.dist = None
        self.flush = None

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        if self.dist is None:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def flush_l2(self):
        self.flush.zero_()  # 256 MiB write > 126 MB L2
        self.torch.cuda.synchronize()


class StatusGather:
    """The path's only exchange: every rank contributes the status bytes of its shard.  Persistent pinned + device
    buffers, asynchronous copy, one all_gather_into_tensor per step on the device, no host read inside the loop."""

    def __init__(self, ctx, n):
        t = ctx.torch
        self.ctx = ctx
        self.pinned = t.empty(n, dtype=t.uint8).pin_memory()
        self.np = self.pinned.numpy()
        if ctx.dist is not None:
            self.dev = t.empty(n, dtype=t.uint8, device="cuda")
            self.all = t.empty(n * ctx.world, dtype=t.uint8, device="cuda")

    def step(self):
        if self.ctx.dist is not None:
            self.dev.copy_(self.pinned, non_blocking=True)
            self.ctx.dist.all_gather_into_tensor(self.all, self.dev)

    def result(self):
        return self.all.cpu().numpy() if self.ctx.dist is not None else self.np


def timed_device_loop(ctx, batch, steps, warmup, stage_fn=None):
    """W untimed + K timed kernel-only passes over a device-resident batch.  Returns (ms list, stage list)."""
    for _ in range(warmup):
        batch.verify(want_status=False)
    ctx.barrier()
    ms_all, stages = [], []
    for _ in range(steps):
        ctx.flush_l2()
        _, ms = batch.verify(want_status=False)
        ms_all.append(ms)
        if stage_fn:
            stages.append(stage_fn())
    ctx.barrier()
    return ms_all, stages


def timed_e2e_loop(ctx, call, gather, steps, warmup):
    """E2E through the public API: host buffers in, status bytes out, exchange included.  Wall clock between barriers,
    max over ranks."""
    for _ in range(warmup):
        call()
    ctx.barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        call()
        if gather is not None:
            gather.step()
    ctx.barrier()
    return ctx.max_over_ranks(time.perf_counter() - t0)


def bench_groth16(ctx, pkg, peak, work):
    args, n, world = ctx.args, ctx.args.batch, ctx.world
    torch = ctx.torch
    vk, proofs, inputs, expected = pkg.groth16_synth(SEED, n, first_index=ctx.rank * n)
    t_proofs = torch.from_numpy(proofs).pin_memory()
    t_inputs = torch.from_numpy(inputs).pin_memory()
    np_proofs, np_inputs = t_proofs.numpy(), t_inputs.numpy()
    h2d, d2h = np_proofs.nbytes + np_inputs.nbytes, n
    batch = pkg.Groth16DeviceBatch(vk, proofs, inputs)

    sampler = ClockSampler(ctx.local_rank)
    for _ in range(args.warmup):
        batch.verify(want_status=False)
    ctx.barrier()
    if ctx.rank == 0:
        sampler.start()
    launches0 = pkg.launch_count()
    t0 = time.perf_counter()
    kernel_ms, split_ms = timed_device_loop(ctx, batch, args.steps, 0, pkg.last_stage_ms)
    wall_kernel = time.perf_counter() - t0
    launches = pkg.launch_count() - launches0
    clocks = sampler.stop() if ctx.rank == 0 else None
    status, _ = batch.verify(want_status=True)
    assert (status == expected).all(), "verdict mismatch against the generator's expected verdicts"
    dev_ms = ctx.max_over_ranks(sum(kernel_ms))  # device time of K steps, max over ranks
    value = world * n * args.steps / (dev_ms * 1e-3)

    gather = StatusGather(ctx, n)
    e2e_s = timed_e2e_loop(ctx, lambda: pkg.Groth16Verifier.verify_batch(np_proofs, vk, np_inputs, out=gather.np), gather,
                           args.steps, max(1, args.warmup // 2))
    assert (gather.np == expected).all()
    allst = gather.result()
    assert allst.shape[0] == world * n and (allst[ctx.rank * n:(ctx.rank + 1) * n] == expected).all()
    e2e_value = world * n * args.steps / e2e_s
    batch.free()
    if ctx.rank != 0:
        return None

    macs_per_proof = work["groth16_macs"]
    step_achieved = macs_per_proof * n * args.steps / (sum(kernel_ms) * 1e-3)
    # stages of a step (CUDA events inside the library): [prepare + Miller, final exponentiation, prepare alone]
    front_ms = sum(s[0] for s in split_ms)
    finish_ms = sum(s[1] for s in split_ms)
    prepare_ms = sum(s[2] for s in split_ms if len(s) > 2)
    multi = finish_ms > 0
    miller_ms = front_ms - prepare_ms
    dom_macs = (work["groth16_miller_macs"] - (work.get("groth16_prepare_macs", 0) if prepare_ms > 0 else 0)) if multi \
        else macs_per_proof
    achieved = dom_macs * n * args.steps / (miller_ms * 1e-3)
    pk = peak["wide_mac_per_s"]
    roofline = {
        "bound": "int32-imad", "achieved": achieved / 1e12, "peak": pk / 1e12, "unit": "TMAC/s", "frac": achieved / pk,
        "traffic": ncu_traffic("k_groth16_miller", n),
        "kernel": "k_groth16_miller" if multi else "k_groth16_verify",
        "kernel_ms_per_launch": miller_ms / args.steps, "macs_per_launch": dom_macs * n,
        "share_of_step": miller_ms / sum(kernel_ms),
        "step": {"achieved": step_achieved / 1e12, "frac": step_achieved / pk,
                 "kernels": ["k_groth16_prepare", "k_groth16_miller", "k_groth16_finish"] if multi else ["k_groth16_verify"],
                 "prepare_ms_per_launch": prepare_ms / args.steps,
                 "prepare_frac": (work.get("groth16_prepare_macs", 0) * n * args.steps / (prepare_ms * 1e-3) / pk)
                 if prepare_ms > 0 else None,
                 "finish_ms_per_launch": finish_ms / args.steps,
                 "finish_frac": (work["groth16_finish_macs"] * n * args.steps / (finish_ms * 1e-3) / pk) if multi else None},
        "macs_per_proof": macs_per_proof, "fp_mul_per_proof": work["groth16_fp_mul"],
        "peak_source": "measured live: bn254v_imad_peak (independent IMAD.WIDE.U32 accumulate chains, 8 warps/SMSP, "
                       "all SMs); MEASURED_PEAKS.json holds no integer peak",
        "peak_imad32_tmacs": peak["lo_mac_per_s"] / 1e12,
        "hbm_gbs_algorithmic": (h2d + d2h) * args.steps / (sum(kernel_ms) * 1e-3) / 1e9,
        "note": "tensor cores unused: carry-chained multiprecision integer arithmetic; HBM traffic negligible",
    }
    return {"value": value, "ms_per_step": dev_ms / args.steps, "roofline": roofline, "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "launches": int(launches), "wall_s_kernel_loop": wall_kernel,
            "cpu_inputs": (vk, proofs, inputs, expected)}


def bench_all_valid(ctx, pkg, peak, work):
    """Opt-in aggregate check (bn254v_groth16_batch_all_valid) on an all-valid batch of the headline size: the valid half
    of 2 x batch trapdoor proofs.  Through the C ABI with host buffers (there is no device-resident form): `value` is
    from the device time of the call (CUDA events inside the library), `e2e` from the wall clock around it."""
    args, n, world = ctx.args, ctx.args.batch, ctx.world
    torch = ctx.torch
    vk, proofs, inputs, expected = pkg.groth16_synth(SEED + 1, 2 * n, first_index=ctx.rank * 2 * n)
    keep = expected == pkg.OK_TRUE
    assert int(keep.sum()) == n
    t_proofs = torch.from_numpy(np.ascontiguousarray(proofs[keep])).pin_memory()
    t_inputs = torch.from_numpy(np.ascontiguousarray(inputs[keep])).pin_memory()
    vp, vi = t_proofs.numpy(), t_inputs.numpy()
    ver = pkg.Groth16Verifier
    for _ in range(max(2, args.warmup)):
        assert ver.batch_all_valid(vp, vk, vi) is True
    ctx.barrier()
    dev_ms, tail_ms = 0.0, 0.0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.flush_l2()
        ok = ver.batch_all_valid(vp, vk, vi)
        ms = pkg.last_stage_ms()
        dev_ms += ms[0] + ms[1]
        tail_ms += ms[1]
        assert ok is True
    ctx.barrier()
    e2e_s = ctx.max_over_ranks(time.perf_counter() - t0)
    dev_ms = ctx.max_over_ranks(dev_ms)
    # one corrupted member: the answer flips
    vp[n // 3] = proofs[np.flatnonzero(~keep)[0]]
    vi[n // 3] = inputs[np.flatnonzero(~keep)[0]]
    assert ver.batch_all_valid(vp, vk, vi) is False
    if ctx.rank != 0:
        return None
    macs = work.get("groth16_agg_macs", 0)
    pk = peak["wide_mac_per_s"]
    achieved = macs * n * args.steps / ((dev_ms - tail_ms) * 1e-3) if macs else None
    return {"metric": "groth16_proofs_checked_per_sec_all_valid_batches", "unit": UNIT,
            "value": world * n * args.steps / (dev_ms * 1e-3), "ms_per_step": dev_ms / args.steps, "steps": args.steps,
            "scaling": "weak",
            "config": {"workload": "2^%d VALID trapdoor Groth16 proofs per GPU (the valid half of 2^%d); opt-in aggregate check "
                                   "'is every proof valid?' (bn254v_groth16_batch_all_valid, scalars drawn by the library); "
                                   "each GPU answers for its own shard" % (n.bit_length() - 1, n.bit_length()),
                       "proofs_per_gpu": n},
            "e2e": {"value": world * n * args.steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": vp.nbytes + vi.nbytes + 16 * n, "d2h_bytes_per_step": n + 1},
            "roofline": {"bound": "int32-imad", "kernel": "k_groth16_agg_prepare + k_groth16_agg_miller (+ product tree)",
                         "achieved": achieved / 1e12 if achieved else None, "peak": pk / 1e12, "unit": "TMAC/s",
                         "frac": achieved / pk if achieved else None, "traffic": ncu_traffic("k_groth16_agg_miller", n),
                         "macs_per_proof": macs,
                         "batch_tail_ms": tail_ms / args.steps},
            "note": "changes semantics (one answer per batch, soundness error <= 2^-126): never used by verify_batch"}


def bench_plonk(ctx, pkg, peak, work):
    """configs[2]: `--plonk-batch` proofs per GPU (4 bundled SP1 proofs replicated, 50 % mutated: half of the mutated
    ones are rejected before the MSMs, the other half in the final pairing check)."""
    import workloads
    args, n, world = ctx.args, ctx.args.plonk_batch, ctx.world
    torch = ctx.torch
    steps, warmup = args.steps, max(2, args.warmup)
    proofs, inputs, rnd, expected = workloads.plonk_workload(n, seed=3 + ctx.rank)
    vk = workloads.plonk_vk_bytes()
    batch = pkg.PlonkDeviceBatch(vk, proofs, inputs, rnd)
    kernel_ms, stages = timed_device_loop(ctx, batch, steps, warmup, pkg.last_stage_ms)
    status, _ = batch.verify(want_status=True)
    assert (status == expected).all(), "PlonK status mismatch"
    dev_ms = ctx.max_over_ranks(sum(kernel_ms))
    value = world * n * steps / (dev_ms * 1e-3)
    tp = [torch.from_numpy(x).pin_memory().numpy() for x in (proofs, inputs, rnd)]
    gather = StatusGather(ctx, n)
    e2e_s = timed_e2e_loop(ctx, lambda: pkg.PlonkVerifier.verify_batch(tp[0], vk, tp[1], rnd=tp[2], out=gather.np), gather,
                           steps, 1)
    assert (gather.np == expected).all()
    batch.free()
    if ctx.rank != 0:
        return None
    pk = peak["wide_mac_per_s"]
    joint = n >= int(os.environ.get("BN254V_PLONK_JOINT_MIN", 1 << 15))  # launch::plonk_joint_min (csrc/k_plonk.cu)
    sm = work["plonk_joint_stage_macs" if joint else "plonk_stage_macs"]
    n_full = int((expected != 6).sum())  # survivors of stage A (6 = ERR_OPENING_POLY_MISMATCH, rejected early)
    names = ["stage_a", "terms0", "stage_c", "terms1", "stage_e"]
    sms = torch.cuda.get_device_properties(ctx.local_rank).multi_processor_count
    trio = n <= int(os.environ.get("BN254V_TRIO_MAX", sms * 128))  # launch::trio_max_items (csrc/k_groth16.cu)
    kern = {"stage_a": "k_plonk_stage_a", "terms0": "k_plonk_terms", "stage_c": "k_plonk_stage_c",
            "terms1": "k_plonk_terms",
            "stage_e": "k_plonk_stage_e3" if trio else "k_plonk_stage_e"}  # (E3: three lanes per proof, after stage D)
    st_ms = [sum(s[i] for s in stages) / steps for i in range(5)]
    per_stage = {}
    for i, nm in enumerate(names):
        units = n if nm == "stage_a" else n_full
        macs = (sm[nm] if nm != "stage_a" else 0) * units
        if nm == "stage_a":  # early rejects stop a little before the end of stage A
            macs = sm["stage_a"] * n_full + work["plonk_early_reject_macs"] * (n - n_full)
        per_stage[nm] = {"kernel": kern[nm], "ms_per_launch": st_ms[i], "macs_per_launch": macs,
                         "frac": macs / (st_ms[i] * 1e-3) / pk if st_ms[i] > 0 else None}
    dom = max(names, key=lambda nm: per_stage[nm]["ms_per_launch"])
    total_macs = sum(per_stage[nm]["macs_per_launch"] for nm in names)
    step_ms = sum(kernel_ms) / steps
    roofline = {"bound": "int32-imad", "kernel": per_stage[dom]["kernel"] + " (" + dom + ")",
                "achieved": per_stage[dom]["macs_per_launch"] / (per_stage[dom]["ms_per_launch"] * 1e-3) / 1e12,
                "peak": pk / 1e12, "unit": "TMAC/s", "frac": per_stage[dom]["frac"], "traffic": ncu_traffic(kern[dom], n),
                "share_of_step": per_stage[dom]["ms_per_launch"] / step_ms,
                "step": {"achieved": total_macs / (step_ms * 1e-3) / 1e12, "frac": total_macs / (step_ms * 1e-3) / pk},
                "stages": per_stage, "macs_full_path": sum(sm.values()),
                "msm_form": "joint (shared doublings)" if joint else "one thread per term",
                "proofs_reaching_the_msms": n_full}
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            ref_cpu = _ref_cpu()
            cores = host_cores()
            sample = min(n, max(64 * cores, 256))
            dtc, stc = ref_cpu.plonk_verify_batch(vk, proofs[:sample], inputs[:sample], rnd[:sample], threads=cores)
            assert (stc == expected[:sample]).all(), "CPU PlonK oracle disagrees"
            cpu = {"value": sample / dtc, "unit": "proofs/s", "cores": cores, "kind": "port",
                   "sample": "first %d records of the same batch, one pass, %.1f s" % (sample, dtc)}
        except Exception as e:
            cpu = {"value": None, "sample": "unavailable: %r" % (e,)}
    return {"metric": "plonk_proofs_verified_per_sec", "unit": "proofs/s", "value": value, "ms_per_step": dev_ms / steps,
            "steps": steps, "scaling": "weak",
            "config": {"workload": "2^%d PlonK proofs per GPU = 4 bundled SP1 fixtures replicated, 50%% mutated "
                                   "(BASELINE.json configs[2]), on-device SHA-256 transcript" % (n.bit_length() - 1),
                       "proofs_per_gpu": n},
            "e2e": {"value": world * n * steps / e2e_s, "unit": "proofs/s",
                    "h2d_bytes_per_step": int(sum(x.nbytes for x in tp)), "d2h_bytes_per_step": n},
            "roofline": roofline, "cpu_baseline": cpu}


def pairing_roofline(work, pk, n, steps, kernel_ms, stages):
    """Big batches run as two launches (k_pairing_miller<4> | k_pairing_finish, stage times from CUDA events inside the
    library); batches of one wave or less as one fused launch (stage [0] is then empty)."""
    macs = work["pairing_product_k4_macs"]
    ach = macs * n * steps / (sum(kernel_ms) * 1e-3)
    out = {"bound": "int32-imad", "achieved": ach / 1e12, "peak": pk / 1e12, "unit": "TMAC/s", "frac": ach / pk,
           "macs_per_set": macs}
    mil = sum(s[0] for s in stages if len(s) >= 2) / max(1, len(stages))
    fin = sum(s[1] for s in stages if len(s) >= 2) / max(1, len(stages))
    if mil > 0.05 * fin:  # two launches
        m_fin = work["groth16_finish_macs"]  # the final exponentiation and the comparison: the same routine
        m_mil = macs - m_fin
        out.update({"kernel": "k_pairing_miller<4>", "kernel_ms_per_launch": mil, "macs_per_launch": m_mil * n,
                    "achieved": m_mil * n / (mil * 1e-3) / 1e12, "share_of_step": mil / (mil + fin),
                    "traffic": ncu_traffic("k_pairing_miller", n)})
        out["frac"] = out["achieved"] * 1e12 / pk
        out["step"] = {"achieved": ach / 1e12, "frac": ach / pk, "kernels": ["k_pairing_miller<4>", "k_pairing_finish"],
                       "finish_ms_per_launch": fin, "finish_frac": m_fin * n / (fin * 1e-3) / pk}
    else:
        out.update({"kernel": "k_pairing_product<4>", "share_of_step": 1.0, "traffic": ncu_traffic("k_pairing_product", n)})
    return out


def bench_pairing(ctx, pkg, peak, work):
    """configs[3]: `--pairing-batch` random 4-pair sets per GPU, all G2 variable; half of the sets multiply to 1."""
    args, n, world = ctx.args, ctx.args.pairing_batch, ctx.world
    torch = ctx.torch
    steps, warmup = max(2, args.steps // 3), 2
    g1, g2, expected = pkg.pairing_synth(11, n, k=4, first_index=ctx.rank * n)
    batch = pkg.PairingDeviceBatch(g1, g2, 4)
    kernel_ms, stages = timed_device_loop(ctx, batch, steps, warmup, pkg.last_stage_ms)
    one, _ = batch.verify(want_status=True)
    assert (one == expected).all(), "pairing product mismatch"
    dev_ms = ctx.max_over_ranks(sum(kernel_ms))
    value = world * n * steps / (dev_ms * 1e-3)
    t1, t2 = torch.from_numpy(g1).pin_memory().numpy(), torch.from_numpy(g2).pin_memory().numpy()
    res = {}

    def call():
        res["one"] = pkg.pairing_product_batch(t1, t2, 4)
    e2e_s = timed_e2e_loop(ctx, call, None, steps, 1)
    assert (res["one"] == expected).all()
    batch.free()
    if ctx.rank != 0:
        return None
    pk = peak["wide_mac_per_s"]
    macs = work["pairing_product_k4_macs"]
    ach = macs * n * steps / (sum(kernel_ms) * 1e-3)
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        try:
            ref_cpu = _ref_cpu()
            cores = host_cores()
            sample = min(n, max(64 * cores, 256))
            dtc, one_c, _, _ = ref_cpu.pairing_product_batch(g1[:sample], g2[:sample], 4, threads=cores)
            assert (one_c == expected[:sample]).all()
            cpu = {"value": sample / dtc, "unit": "4-pair sets/s", "cores": cores, "kind": "port",
                   "sample": "first %d sets of the same batch, one pass, %.1f s" % (sample, dtc)}
        except Exception as e:
            cpu = {"value": None, "sample": "unavailable: %r" % (e,)}
    return {"metric": "pairing_product_sets_per_sec", "unit": "4-pair sets/s", "value": value,
            "pair_miller_loops_per_sec": 4 * value, "final_exponentiations_per_sec": value,
            "ms_per_step": dev_ms / steps, "steps": steps, "scaling": "weak",
            "config": {"workload": "2^%d random 4-pair sets per GPU, all G2 variable, half of them multiply to 1 "
                                   "(BASELINE.json configs[3])" % (n.bit_length() - 1), "sets_per_gpu": n},
            "e2e": {"value": world * n * steps / e2e_s, "unit": "4-pair sets/s",
                    "h2d_bytes_per_step": int(t1.nbytes + t2.nbytes), "d2h_bytes_per_step": n},
            "roofline": pairing_roofline(work, pk, n, steps, kernel_ms, stages),
            "cpu_baseline": cpu}


def bench_mixed(ctx, pkg, peak, work):
    """configs[4]: `--mixed-total` items over all GPUs (half Groth16, half PlonK, interleaved one to one), sharded by
    index: rank r owns items [r, r + 1) * total / N.  Strong scaling.  Device-timed value: the two device-resident
    batches of the shard back to back; e2e: one bn254v_verify_many call over the interleaved host items."""
    import numpy as np
    import workloads
    args, world = ctx.args, ctx.world
    total = args.mixed_total
    per_rank = total // world
    half = per_rank // 2
    steps, warmup = max(2, args.steps // 5), 1
    vk_g, pr_g, in_g, exp_g = pkg.groth16_synth(SEED + 1, half, first_index=ctx.rank * half)
    pr_p, in_p, rnd_p, exp_p = workloads.plonk_workload(half, seed=100 + ctx.rank)
    vk_p = workloads.plonk_vk_bytes()
    bg = pkg.Groth16DeviceBatch(vk_g, pr_g, in_g)
    bp = pkg.PlonkDeviceBatch(vk_p, pr_p, in_p, rnd_p)

    class Both:
        def verify(self, want_status=False):
            s1, m1 = bg.verify(want_status)
            s2, m2 = bp.verify(want_status)
            return (s1, s2), m1 + m2
    kernel_ms, _ = timed_device_loop(ctx, Both(), steps, warmup)
    (s1, s2), _ = Both().verify(True)
    assert (s1 == exp_g).all() and (s2 == exp_p).all()
    dev_ms = ctx.max_over_ranks(sum(kernel_ms))
    value = world * 2 * half * steps / (dev_ms * 1e-3)
    bg.free(), bp.free()
    gather = StatusGather(ctx, 2 * half)
    # the bn254v_item array (pointers into the host arrays above) is the caller's input and is built once; each timed
    # call gathers the records, copies them to the device, verifies and copies the status bytes back
    items = pkg.MixedItems(vk_g, pr_g, in_g, vk_p, pr_p, in_p, rnd_p)
    e2e_s = timed_e2e_loop(ctx, lambda: items.verify(out=gather.np), gather, steps, 1)
    assert (gather.np[0::2] == exp_g).all() and (gather.np[1::2] == exp_p).all()
    if ctx.rank != 0:
        return None
    pk = peak["wide_mac_per_s"]
    n_full = int((exp_p != 6).sum())
    pj = half >= int(os.environ.get("BN254V_PLONK_JOINT_MIN", 1 << 15))  # chunks of 2^16 proofs: the joint MSM form
    macs = work["groth16_macs"] * half + work["plonk_joint_full_path_macs" if pj else "plonk_full_path_macs"] * n_full + \
        work["plonk_early_reject_macs"] * (half - n_full)
    ach = macs * steps / (sum(kernel_ms) * 1e-3)
    return {"metric": "mixed_items_verified_per_sec", "unit": "items/s", "value": value, "ms_per_step": dev_ms / steps,
            "steps": steps, "scaling": "strong",
            "config": {"workload": "2^%d items over all GPUs: Groth16 (as configs[1]) and PlonK (as configs[2]) "
                                   "interleaved one to one (BASELINE.json configs[4]), sharded by index" %
                                   (total.bit_length() - 1), "items_total": world * 2 * half, "items_per_gpu": 2 * half},
            "e2e": {"value": world * 2 * half * steps / e2e_s, "unit": "items/s", "api": "bn254v_verify_many over a prebuilt bn254v_item array",
                    "h2d_bytes_per_step": int(pr_g.nbytes + in_g.nbytes + pr_p.nbytes + in_p.nbytes + rnd_p.nbytes),
                    "d2h_bytes_per_step": 2 * half},
            "roofline": {"bound": "int32-imad", "kernel": "k_groth16_miller + k_groth16_finish + k_plonk_*",
                         "achieved": ach / 1e12, "peak": pk / 1e12, "unit": "TMAC/s", "frac": ach / pk, "traffic": None}}


def bench_single_process(args, pkg, world):
    """Rank 0 alone drives `world` GPUs through ONE library call: bn254v_init over all devices and the in-library
    sharding loop of bn254v.cu (contiguous index ranges, one stream per device).  N x 2^17 proofs per step, i.e. 2^20
    proofs in one step at N = 8.  Runs before the other ranks touch their GPUs (they wait in the rendezvous)."""
    import torch
    n = world * (1 << 17)
    pkg.init(list(range(world)))
    assert pkg.load_library().bn254v_device_count() == world
    vk, proofs, inputs, expected = pkg.groth16_synth(SEED + 7, n)
    batch = pkg.Groth16DeviceBatch(vk, proofs, inputs)
    for _ in range(2):
        batch.verify(want_status=False)
    ms = [batch.verify(want_status=False)[1] for _ in range(4)]
    status, _ = batch.verify(want_status=True)
    assert (status == expected).all()
    tp, ti = torch.from_numpy(proofs).pin_memory().numpy(), torch.from_numpy(inputs).pin_memory().numpy()
    out = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
    pkg.Groth16Verifier.verify_batch(tp, vk, ti, out=out)  # warm-up: scratch buffers of this shape
    t0 = time.perf_counter()
    for _ in range(3):
        pkg.Groth16Verifier.verify_batch(tp, vk, ti, out=out)
    e2e = 3 * n / (time.perf_counter() - t0)
    assert (out == expected).all()
    batch.free()
    pkg.shutdown()
    return {"devices": world, "proofs_per_step": n, "ms_per_step": sum(ms) / len(ms),
            "value": n * len(ms) / (sum(ms) * 1e-3), "unit": UNIT, "e2e_value": e2e, "verdicts": "all as expected",
            "note": "one process, bn254v_init over all devices; device time = max over devices (CUDA events)"}


def run_b200(args):
    import torch
    import __graft_entry__ as ge

    ctx = Ctx(args)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback")
    wl = set(args.workloads.split(",")) if not args.no_secondary else set()
    pkg = ge.load_package()
    single = None
    if ctx.world > 1 and "single" in wl and ctx.rank == 0 and torch.cuda.device_count() >= ctx.world:
        single = bench_single_process(args, pkg, ctx.world)
    torch.cuda.set_device(ctx.local_rank)
    if ctx.world > 1:
        import datetime
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", ctx.local_rank),
                                timeout=datetime.timedelta(minutes=20))
        ctx.dist = dist
    pkg.init([ctx.local_rank])
    ctx.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    peak = pkg.imad_peak(2048)
    work = work_per_proof()

    g = bench_groth16(ctx, pkg, peak, work)
    extra = {}
    if "plonk" in wl:
        extra["plonk"] = bench_plonk(ctx, pkg, peak, work)
    if "pairing" in wl:
        extra["pairing"] = bench_pairing(ctx, pkg, peak, work)
    if "mixed" in wl:
        extra["mixed"] = bench_mixed(ctx, pkg, peak, work)
    if "allvalid" in wl:
        extra["all_valid"] = bench_all_valid(ctx, pkg, peak, work)
    total_launches = pkg.launch_count()
    if ctx.rank != 0:
        if ctx.dist is not None:
            ctx.dist.destroy_process_group()
        return

    n, world = args.batch, ctx.world
    cpu = None
    if not args.no_cpu_baseline and world == 1:  # reported at N=1 only
        try:
            vk, proofs, inputs, expected = g["cpu_inputs"]
            cores = host_cores()
            sample = min(args.cpu_sample or max(cores * 512, 512), n)
            dt, st_cpu = _ref_cpu().groth16_verify_batch(vk, proofs[:sample], inputs[:sample], cores)
            assert (st_cpu == expected[:sample]).all(), "CPU oracle disagrees with expected verdicts"
            cpu = {"value": sample / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": "first %d proofs of the same batch, one pass, %.1f s" % (sample, dt)}
        except Exception as e:  # the baseline is a reported number, never a dependency of the product path
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "unavailable: %r" % (e,)}
    if single is not None:
        extra["single_process"] = single
    line = {
        "metric": METRIC, "value": g["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": g["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32", "data": "synthetic",
        "config": {"workload": "2^%d trapdoor-simulated Groth16 proofs per GPU, 2 public inputs, 50%% corrupted "
                               "(BASELINE.json configs[1])" % (n.bit_length() - 1),
                   "proofs_per_gpu": n, "seed": SEED, "l2": "flushed between timed iterations (256 MiB write)",
                   "parallelism": "proof-index sharding, %d rank(s), no data-path collective" % world},
        "roofline": g["roofline"], "cpu_baseline": cpu, "clocks": g["clocks"], "e2e": g["e2e"],
        "gpu_launches": g["launches"], "gpu_launches_whole_run": int(total_launches),
        "pairings_per_sec": 3 * g["value"], "wall_s_kernel_loop": g["wall_s_kernel_loop"],
        **extra,
    }
    print(json.dumps(line))
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
