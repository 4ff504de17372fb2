#!/usr/bin/env python3
"""bench.py -- headline benchmark: KZG-commit MSM, BN254 G1, 2^20 (scalar, point) pairs per GPU.

    python bench.py --gpus N --steps K --warmup W            # this framework (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU restatement of the reference path

One "step" = one `msm_bigint(&powers_of_g, scalars)` of 2^20 pairs per GPU (BASELINE.json
configs[1], the configuration `metric` "MSM 2^20 G1 ms" is quoted on).  The commit key
([beta^i]G, known beta) is resident, as `ProvingKey.commit_key` is across proofs.

  value   device-timed: scalars already in HBM, result left in HBM as one XYZZ point
          (CUDA events on the launching stream, max over ranks).
  e2e     the same MSM through the C-ABI call a user makes (`jf_msm`): scalars in pinned host
          memory, H2D copy + kernels + D2H of the result + host normalisation inside the timed
          region; for N > 1 it includes the NCCL all-gather of the per-GPU partial sums and
          the final combine.
  N > 1   weak scaling: the job is ONE MSM of N * 2^20 pairs, sharded by point range (rank r
          holds key[r*2^20, (r+1)*2^20) and the matching scalars); one 128-byte-per-rank NCCL
          all-gather joins the partial sums.  `value` is the time of that whole step.

Inputs are larger than L2: the step rotates through 8 distinct scalar vectors (256 MB) and
gathers from 1 GiB of window tables.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np

METRIC = "MSM 2^20 G1 ms (BN254, KZG commit)"
LOG_N = 20
N_SETS = 8
BETA = 0x1D3C7A5B9E8F60412B7A6C5D4E3F20198A7B6C5D4E3F2A1B0C9D8E7F6A5B4C3  # fixed, known
SEED = 0x6A656C6C79666973


def _clock_sampler(stop, rows, dev):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", "-i", str(dev), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
    except Exception:
        return
    try:
        while not stop.is_set():
            line = p.stdout.readline()
            if not line:
                break
            rows.append([c.strip() for c in line.split(",")])
    finally:
        p.terminate()


def _clock_summary(rows):
    sm, mx, reasons = [], [], set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for r in rows:
        try:
            sm.append(float(r[0]))
            mx.append(float(r[1]))
        except Exception:
            continue
        for name, v in zip(names, r[3:7]):
            if v.lower().startswith("active"):
                reasons.add(name)
    if not sm:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def run_reference(args):
    """CPU arm: the oracle's restatement of ark-ec 0.4.2 msm_bigint (signed-digit Pippenger,
    parallel across windows only, like rayon in the reference) on the host cores.  The Rust
    reference cannot be compiled in this image, so kind = "port" ("Oracle")."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import coracle as co
    co.build()
    threads = co.max_threads()
    windows = co.msm_windows(1 << LOG_N, 254)
    # bounded sample: find a size whose (steps + warmup) MSMs fit in ~150 s
    n0 = 1 << 14
    ks = co.random_field_elems("bn254_fr", n0, SEED, False)
    pts0 = co.fixed_base_mul("bn254", ks)
    s0 = co.random_field_elems("bn254_fr", n0, SEED + 1, False)
    t = time.perf_counter()
    co.msm("bn254", pts0, s0)
    t14 = time.perf_counter() - t
    budget = 150.0 / max(args.steps + args.warmup, 1)
    log_s = 14
    while log_s < LOG_N and t14 * (2 ** (log_s + 1 - 14)) * 0.8 < budget:
        log_s += 1
    n = 1 << log_s
    # points: tile the 2^14 known points (the bucket method's cost does not depend on which
    # points they are), scalars: fresh uniform values
    pts = np.ascontiguousarray(np.tile(pts0, (n // n0, 1)))
    times = []
    for i in range(args.warmup + args.steps):
        s = co.random_field_elems("bn254_fr", n, SEED + 10 + i, False)
        t = time.perf_counter()
        co.msm("bn254", pts, s)
        dt = time.perf_counter() - t
        if i >= args.warmup:
            times.append(dt)
    ms_sample = 1e3 * float(np.mean(times))
    scale = (1 << LOG_N) / n
    ms = ms_sample * scale
    sample = ("%d MSM(s) of 2^%d pairs per step, %d threads (ark-ec parallelises over its %d windows only)%s"
              % (1, log_s, threads, co.msm_windows(n, 254),
                 "" if log_s == LOG_N else "; scaled linearly x%d to 2^%d" % (int(scale), LOG_N)))
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": "ms", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64x4 Montgomery (CPU, __int128)", "data": "synthetic",
        "config": {"workload": "KZG commit MSM, BN254 G1, N=2^20 uniform canonical scalars, CPU restatement of "
                               "ark-ec msm_bigint", "windows_at_2^20": windows},
        "cpu_baseline": {"value": ms, "unit": "ms", "cores": min(threads, co.msm_windows(n, 254)), "kind": "port",
                         "sample": sample},
        "e2e": {"value": ms, "unit": "ms", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(line)



# ------------------------------------------------------------------------------------------------
NTT_FIELD, NTT_LOG_N, NTT_BATCH = "bn254_fr", 22, 16


def ntt_leg(ctx, co, torch, args, world, rank, barrier, max_over_ranks, stream):
    """Second metric of BASELINE.json: "NTT 2^22 Melem/s".  One step = one batched forward COSET NTT
    (`coset.fft`, prover.rs:552-567) of 16 polynomials of 2^22 BN254-Fr coefficients per GPU, in place,
    natural order in and out.  N > 1: the batch is sharded by polynomial (16 per rank, no collective)."""
    n, batch = 1 << NTT_LOG_N, NTT_BATCH
    distinct = 4
    host = co.random_field_elems(NTT_FIELD, n * distinct, SEED + 77 + rank, True).reshape(distinct, n, 4)
    pinned = torch.empty((batch, n, 4), dtype=torch.int64).pin_memory()
    for b in range(batch):
        pinned[b].copy_(torch.from_numpy(host[b % distinct].view(np.int64)))
    d = torch.empty((batch, n, 4), dtype=torch.int64, device="cuda")
    d.copy_(pinned)
    off = co.field_op(NTT_FIELD, "to_mont", np.array([[5, 0, 0, 0]], dtype=np.uint64))[0]  # Fr::GENERATOR
    # correctness guard on this very configuration: spot-check out[i] = p(g w^i) by Horner (oracle) for poly 0
    ctx.ntt_device(NTT_FIELD, d.data_ptr(), NTT_LOG_N, False, off, batch=batch)
    torch.cuda.synchronize()
    if rank == 0:
        import pyref
        F = pyref.BN254_FR
        got = d[0].cpu().numpy().view(np.uint64)
        w = pow(F.two_adic_root, 1 << (F.two_adicity - NTT_LOG_N), F.p)
        for i in (0, 1, n // 3, n - 1):
            xpt = (5 * pow(w, i, F.p)) % F.p
            want = co.poly_eval(NTT_FIELD, host[0], co.ints_to_limbs([F.to_mont(xpt)], 4)[0])
            if not np.array_equal(got[i], want):
                raise SystemExit("bench.py: coset NTT output does not match Horner evaluation; refusing to time it")
    steps = args.steps
    for i in range(args.warmup):
        ctx.ntt_device(NTT_FIELD, d.data_ptr(), NTT_LOG_N, False, off, batch=batch)
    barrier()
    l0 = ctx.launch_count
    ctx.profile(True, dominant_only=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(steps):
        ctx.ntt_device(NTT_FIELD, d.data_ptr(), NTT_LOG_N, False, off, batch=batch)
    e1.record(stream)
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / steps)
    prof = ctx.profile_collect()
    ctx.profile(False)
    launches = ctx.launch_count - l0
    # config 3's field (BLS12-381 Fr, GENERATOR 7) on the same buffers, device-timed only
    off_bls = np.array([0x0000000efffffff1, 0x17e363d300189c0f, 0xff9c57876f8457b0, 0x351332208fc5a8c4], dtype=np.uint64)  # 7 R mod r
    for i in range(3):
        ctx.ntt_device("bls12_381_fr", d.data_ptr(), NTT_LOG_N, False, off_bls, batch=batch)
    barrier()
    b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b0.record(stream)
    for i in range(steps):
        ctx.ntt_device("bls12_381_fr", d.data_ptr(), NTT_LOG_N, False, off_bls, batch=batch)
    b1.record(stream)
    barrier()
    ms_bls = max_over_ranks(b0.elapsed_time(b1) / steps)
    # end to end through jf_ntt: pinned host coefficients in, evaluations back in the same host buffer
    e_steps = max(1, min(steps, 4))
    arr = pinned.numpy().view(np.uint64)
    ctx.ntt(NTT_FIELD, arr, NTT_LOG_N, False, off)
    barrier()
    t0 = time.perf_counter()
    for i in range(e_steps):
        ctx.ntt(NTT_FIELD, arr, NTT_LOG_N, False, off)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / e_steps)
    if rank != 0:
        return None
    cnt, tot = prof.get("ntt_pass", (0, 0.0))
    passes = cnt // max(steps, 1)
    pass_ms = tot / max(cnt, 1)
    mul_rate = ctx.microbench(1)
    elems = n * batch
    muls = elems * (NTT_LOG_N / 2.0 + 1.0)          # (n/2) log2 n butterflies + n coset scalings
    peaks = _peaks()
    cpu = None
    if not args.no_cpu:
        x = host[0].copy()
        co.ntt(NTT_FIELD, x, NTT_LOG_N, False, off)   # warm the oracle's twiddle setup
        t0 = time.perf_counter()
        co.ntt(NTT_FIELD, x, NTT_LOG_N, False, off)
        dt = time.perf_counter() - t0
        cpu = {"value": n / dt / 1e6, "unit": "Melem/s", "cores": co.max_threads(), "kind": "port",
               "sample": "1 coset NTT of 2^%d BN254 Fr elements, OpenMP radix-2 restatement of ark-poly" % NTT_LOG_N}
    return {
        "metric": "NTT 2^22 Melem/s (BN254 Fr, forward coset, batch 16 per GPU, in place, natural order)",
        "value": elems * world / (ms * 1e-3) / 1e6, "unit": "Melem/s", "ms_per_step": ms, "higher_is_better": True,
        "passes_per_transform": passes, "gpu_launches": int(launches),
        "bls12_381_fr_melem_per_s": elems * world / (ms_bls * 1e-3) / 1e6,
        "roofline": {"bound": "hbm", "kernel": "ntt_pass", "achieved": 64.0 * elems / (pass_ms * 1e-3) / 1e9,
                     "peak": peaks[0], "unit": "GB/s", "frac": 64.0 * elems / (pass_ms * 1e-3) / 1e9 / peaks[0],
                     "traffic": (_traffic("ntt_pass_bytes_per_element") or 0) * elems or None, "peak_source": peaks[1],
                     "note": "per launch = one Stockham pass over the batch (read 32 B + write 32 B per element); a "
                             "transform is %d passes, so the whole-transform figure is 1/%d of this" % (passes, max(passes, 1))},
        "roofline_int": {"bound": "imad", "achieved": muls / (ms * 1e-3), "peak": mul_rate, "unit": "Montgomery mul/s",
                         "frac": muls / (ms * 1e-3) / mul_rate,
                         "peak_source": "jf_microbench(1): dependent 256-bit Montgomery products in registers, this run",
                         "note": "algorithmic multiplications only ((n/2) log n + n); inter-pass twiddles are overhead"},
        "cpu_baseline": cpu,
        "e2e": {"value": elems * world / (e2e_ms * 1e-3) / 1e6, "unit": "Melem/s", "ms_per_step": e2e_ms,
                "h2d_bytes_per_step": 32 * elems, "d2h_bytes_per_step": 32 * elems},
    }


# ------------------------------------------------------------------------------------------------
PROVE_LOG_N = 20


def prove_leg(ctx, co, torch, args, world, rank, barrier, max_over_ranks, stream, cpu_msm_ms, cpu_ntt_melems):
    """First metric of BASELINE.json: "BN254 2^20-gate PLONK prove ms" (configs[3]).  One step = one
    `PlonkKzgSnark::prove` of the reference's own bench circuit (plonk/benches/bench.rs:29-46, 2^20 gates,
    TurboPlonk, SolidityTranscript) through the C-ABI call `jf_plonk_prove`: witness in host memory
    (32 MiB H2D), proof (13 commitments + 10 evaluations) back on the host, the Fiat-Shamir transcript
    on the host in between.  The call is end to end by construction, so value == e2e.  N > 1: every rank
    proves its own instance (replicas; round 3 needs all polynomials on one GPU)."""
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_circuit as B
    n = 1 << PROVE_LOG_N
    t0 = time.time()
    arr = B.bench_circuit_arrays(ctx, PROVE_LOG_N)
    key = ctx.generate_srs_for_testing("bn254", BETA % co_modulus(), n + 3)
    t_setup = time.time() - t0
    rng = np.random.default_rng(SEED % (1 << 32) + rank)
    bl = rng.integers(0, 1 << 60, size=(17, 4), dtype=np.uint64)      # < 2^252: valid field elements
    wit = torch.from_numpy(arr["witness"].view(np.int64)).pin_memory().numpy().view(np.uint64)
    out = {}
    steps = max(1, min(args.steps, 5))
    for cache, skip, full in ((False, False, False), (False, True, False), (True, True, False), (False, False, True)):
        t0 = time.time()
        pk = jf_mod().PlonkKzgSnark.preprocess(ctx, key, arr["selectors"], arr["sigmas"], arr["k"], arr["wire_vars"],
                                               arr["num_vars"], [], cache_coset_evals=cache, skip_zero_selectors=skip,
                                               full_quotient_coset=full)
        t_pre = time.time() - t0
        proof = jf_mod().PlonkKzgSnark.prove(pk, wit, bl, "solidity")
        if rank == 0 and not cache and not skip and not full:
            # checker: the restated jellyfish verifier (known-beta G1 form) must accept the proof
            sys.path.insert(0, os.path.join(ROOT, "tests"))
            import plonk_ref as P
            import plonk_util as U
            import pyref
            cv = pyref.BN254
            vk = U.vk_from_product(co, cv, pk, B.BN254_K)
            if not P.verify(cv, vk, [], U.proof_to_oracle(co, cv, proof), BETA % co_modulus(), "solidity"):
                raise SystemExit("bench.py: the 2^20 proof is rejected by the restated verifier; refusing to time it")
        for _ in range(2):
            jf_mod().PlonkKzgSnark.prove(pk, wit, bl, "solidity")
        barrier()
        l0 = ctx.launch_count
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        t0 = time.perf_counter()
        for _ in range(steps):
            jf_mod().PlonkKzgSnark.prove(pk, wit, bl, "solidity")
        wall = (time.perf_counter() - t0) * 1e3 / steps
        e1.record(stream)
        barrier()
        dev = e0.elapsed_time(e1) / steps
        launches = (ctx.launch_count - l0) // steps
        ctx.profile(True)
        jf_mod().PlonkKzgSnark.prove(pk, wit, bl, "solidity")
        prof = ctx.profile_collect()
        ctx.profile(False)
        out[(cache, skip, full)] = {"wall_ms": max_over_ranks(wall), "device_ms": max_over_ranks(dev), "launches": int(launches),
                      "preprocess_s": round(t_pre, 3),
                      "kernels_ms": {k: round(v[1], 3) for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:8]}}
        pk.free()
    key.free()
    # BASELINE configs[0] shape (2^16 gates, the size the reference's own bench can run): one data point
    arr16 = B.bench_circuit_arrays(ctx, 16)
    key16 = ctx.generate_srs_for_testing("bn254", BETA % co_modulus(), (1 << 16) + 3)
    pk16 = jf_mod().PlonkKzgSnark.preprocess(ctx, key16, arr16["selectors"], arr16["sigmas"], arr16["k"], arr16["wire_vars"],
                                             arr16["num_vars"], [])
    for _ in range(2):
        jf_mod().PlonkKzgSnark.prove(pk16, arr16["witness"], bl, "solidity")
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        jf_mod().PlonkKzgSnark.prove(pk16, arr16["witness"], bl, "solidity")
    prove16_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)
    pk16.free()
    key16.free()
    if rank != 0:
        return None
    cpu = None
    if cpu_msm_ms is not None and cpu_ntt_melems is not None:
        est = 13 * cpu_msm_ms + (26 * 8 * n + 7 * n) / (cpu_ntt_melems * 1e6) * 1e3
        cpu = {"value": est, "unit": "ms", "cores": co.max_threads(), "kind": "port",
               "sample": "component sum of the CPU restatement timed in this run: 13 MSM(2^20) + 26 coset NTT(2^23) + "
                         "7 iNTT(2^20) (App. A workload); pointwise / Horner / division terms not included. The reference's "
                         "only published figure extrapolates to ~24 s on a 5900X (bench.md:17, 2^15 gates x 32)"}
    base = out[(False, False, False)]
    alt = lambda k, note: {"value": out[k]["wall_ms"], "unit": "ms", "device_ms": out[k]["device_ms"],  # noqa: E731
                           "gpu_launches": out[k]["launches"], "note": note}
    return {
        "metric": "BN254 2^20-gate TurboPlonk prove ms (bench.rs circuit, SolidityTranscript, proof accepted by the restated verifier)",
        "value": base["wall_ms"] / 1.0, "unit": "ms", "ms_per_step": base["wall_ms"], "higher_is_better": False,
        "proofs_per_s": world * 1e3 / base["wall_ms"], "device_ms": base["device_ms"], "gpu_launches": base["launches"],
        "kernels_ms_per_proof": base["kernels_ms"], "setup_s": round(t_setup, 2), "preprocess_s": base["preprocess_s"],
        "with_full_8n_quotient_coset": alt((False, False, True), "round 3 in the reference's literal form: one 8n-point coset NTT per "
                                           "polynomial (prover.rs:552-567) instead of six sub-cosets of n points; same proof bytes"),
        "with_zero_selector_skip": alt((False, True, False), "selector columns that are identically zero (9 of 13 in this circuit: "
                                       "q_lc2-3, q_mul, q_hash, q_ecc) are recognised at preprocess; their coset NTTs and "
                                       "quotient terms are skipped; same proof bytes (tests/test_gpu_plonk.py)"),
        "with_cached_selector_sigma_coset_evals": alt((True, True, False), "additionally the selector / sigma coset evaluations stay "
                                                      "resident (+3.4 GiB): only 7 polynomials are transformed per proof; same proof bytes"),
        "prove_2^16_gates_ms": prove16_ms,
        "cpu_baseline": cpu,
        "e2e": {"value": base["wall_ms"], "unit": "ms", "h2d_bytes_per_step": int(arr["witness"].nbytes + 17 * 32),
                "d2h_bytes_per_step": 13 * 128 + 10 * 32},
        "scaling": "replicas" if world > 1 else "n/a",
    }


def jf_mod():
    import mpc_jellyfish_b200 as jf
    return jf


# ------------------------------------------------------------------------------------------------
def mpc_leg(ctx, co, torch, args, world, rank, barrier, max_over_ranks, stream):
    """BASELINE config 5 shape: one party's local share-wise work in the collaborative prover at n = 2^18
    (plonk/src/multiprover/primitives/multiprover_kzg.rs:128-143, multiprover/proof_system/prover.rs:373-388):
    `MultiproverKZG::batch_commit` of the 5 authenticated wire polynomials (share + MAC vectors = 10 MSMs of
    n + 2 scalars through jf_msm_batch, host buffers) and the share-wise coset NTT of the same 5 polynomials to
    the 8n domain (10 vectors, device resident).  N > 1: one party per GPU (no collective on this path)."""
    n = 1 << 18
    rng = np.random.default_rng(7 + rank)
    key = ctx.generate_srs_for_testing("bn254", BETA % co_modulus(), n + 3)
    pp = jf_mod().UnivariateProverParam(key)
    polys = [jf_mod().AuthenticatedDensePoly(rng.integers(0, 1 << 60, size=(n + 2, 4), dtype=np.uint64),
                                             rng.integers(0, 1 << 60, size=(n + 2, 4), dtype=np.uint64)) for _ in range(5)]
    steps = max(1, min(args.steps, 5))

    def timed_commit(ps):
        for _ in range(2):
            jf_mod().MultiproverKZG.batch_commit(pp, ps)
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            jf_mod().MultiproverKZG.batch_commit(pp, ps)
        return max_over_ranks((time.perf_counter() - t0) * 1e3 / steps)

    pageable_ms = timed_commit(polys)
    # the same vectors in page-locked memory (what jf_host_alloc hands the shim crate for coefficient vectors)
    pin = lambda a: torch.from_numpy(a.view(np.int64)).pin_memory().numpy().view(np.uint64)  # noqa: E731
    pinned_polys = [jf_mod().AuthenticatedDensePoly(pin(pl.share), pin(pl.mac)) for pl in polys]
    commit_ms = timed_commit(pinned_polys)
    m = 8 * n
    d = torch.zeros((10, m, 4), dtype=torch.int64, device="cuda")
    for i, pl in enumerate(polys):
        d[2 * i, : n + 2].copy_(torch.from_numpy(pl.share.view(np.int64)))
        d[2 * i + 1, : n + 2].copy_(torch.from_numpy(pl.mac.view(np.int64)))
    base = d.clone()
    off = np.array([0x1b0d0ef99fffffe6, 0xeaba68a3a32a913f, 0x47d8eb76d8dd0689, 0x15d0085520f5bbc3], dtype=np.uint64)  # 5 R mod r
    for _ in range(2):
        d.copy_(base)
        ctx.ntt_device("bn254_fr", d.data_ptr(), 21, False, off, in_len=n + 2, batch=10)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    tot = 0.0
    for _ in range(steps):
        d.copy_(base)
        e0.record(stream)
        ctx.ntt_device("bn254_fr", d.data_ptr(), 21, False, off, in_len=n + 2, batch=10)
        e1.record(stream)
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ntt_ms = max_over_ranks(tot / steps)
    key.free()
    if rank != 0:
        return None
    return {"workload": "one party, n = 2^18: batch_commit of 5 authenticated wire polynomials (10 MSMs, host scalars) and "
                        "their share-wise coset NTT to 8n (10 vectors, device resident)",
            "batch_commit_ms": commit_ms, "batch_commit_pageable_host_ms": pageable_ms, "msm_per_s": 10 * world / (commit_ms * 1e-3),
            "note": "batch_commit_ms: scalars in page-locked host memory; pageable: plain numpy arrays (the driver stages them)",
            "sharewise_coset_ntt_ms": ntt_ms, "ntt_melem_per_s": 10 * m * world / (ntt_ms * 1e-3) / 1e6,
            "parties": world}


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def _traffic(kernel):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
def run_cuda(args):
    import torch
    import torch.distributed as dist
    import mpc_jellyfish_b200 as jf
    import coracle as co  # input generation + the cpu_baseline leg only

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 1 << LOG_N
    ctx = jf.Context(local)
    # a real (non-default) stream shared by torch (copies, NCCL, events) and the library's kernels
    stream = torch.cuda.Stream(priority=-1)  # high priority: the library's side streams only fill its idle slots
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    # commit key slice of this rank: [beta^(rank*n + i)] G
    t0 = time.time()
    key = ctx.generate_srs_for_testing("bn254", BETA, n, first_power=rank * n)
    t_key = time.time() - t0

    # scalars: N_SETS distinct uniform vectors per rank, canonical BigInts like `into_bigint` yields
    host_sets = []
    for k in range(N_SETS):
        host_sets.append(co.random_field_elems("bn254_fr", n, SEED + 1000 * rank + k, False))
    d_sets = torch.empty((N_SETS, n, 4), dtype=torch.int64, device="cuda")
    pinned = torch.empty((N_SETS, n, 4), dtype=torch.int64).pin_memory()
    for k in range(N_SETS):
        pinned[k].copy_(torch.from_numpy(host_sets[k].view(np.int64)))
    d_sets.copy_(pinned, non_blocking=False)
    d_out = torch.zeros((16,), dtype=torch.int64, device="cuda")          # one XYZZ point (128 B)
    gathered = torch.zeros((world, 16), dtype=torch.int64, device="cuda")
    res_host = torch.zeros((world, 16), dtype=torch.int64).pin_memory()

    def step_device(i):
        k = i % N_SETS
        ctx.msm_device(key, d_sets[k].data_ptr(), n, d_out.data_ptr())
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(-1), d_out)

    def step_e2e(i):
        k = i % N_SETS
        if world == 1:
            # the C-ABI call a user makes: host scalars in, affine point out
            return ctx.msm(key, pinned[k].numpy().view(np.uint64))
        d_sets[k].copy_(pinned[k], non_blocking=True)
        ctx.msm_device(key, d_sets[k].data_ptr(), n, d_out.data_ptr())
        dist.all_gather_into_tensor(gathered.view(-1), d_out)
        res_host.copy_(gathered, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return ctx.msm_combine("bn254", res_host.numpy().view(np.uint64))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- correctness guard: the commitment equals p(beta') * G for the known beta ----------------
    xy, inf = step_e2e(0)
    if rank == 0 and world == 1:
        ev = co.poly_eval("bn254_fr", co.field_op("bn254_fr", "to_mont", host_sets[0]),
                          co.field_op("bn254_fr", "to_mont", co.ints_to_limbs([BETA % co_modulus()], 4))[0])
        want = co.fixed_base_mul("bn254", co.field_op("bn254_fr", "from_mont", ev[None, :]))[0]
        if inf or not np.array_equal(xy, want):
            raise SystemExit("bench.py: MSM result does not match the known-beta identity; refusing to time a wrong kernel")

    # ---- device-timed region ---------------------------------------------------------------------
    for i in range(args.warmup):
        step_device(i)
    barrier()
    stop, rows = threading.Event(), []
    sampler = threading.Thread(target=_clock_sampler, args=(stop, rows, local), daemon=True)
    sampler.start()
    time.sleep(0.25)
    launches0 = ctx.launch_count
    # only the dominant kernel is bracketed by events inside the timed region: events around all ~30 dependent launches
    # of an MSM cost ~0.2 ms per step (the full per-kernel breakdown comes from a separate instrumented pass below)
    ctx.profile(True, dominant_only=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step_device(args.warmup + i)
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    prof_dom = ctx.profile_collect()
    ctx.profile(False)
    launches = ctx.launch_count - launches0
    ms_step = max_over_ranks(ms_total / args.steps)
    # instrumented pass (not timed as a step): every launch bracketed, for the per-kernel breakdown
    ctx.profile(True)
    for i in range(args.steps):
        step_device(args.warmup + i)
    barrier()
    prof = ctx.profile_collect()
    ctx.profile(False)

    # ---- end-to-end region (host buffers, copies inside) -------------------------------------------
    for i in range(min(args.warmup, 3)):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_e2e(args.warmup + i)
    barrier()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / args.steps)
    ntt = None if args.no_ntt else ntt_leg(ctx, co, torch, args, world, rank, barrier, max_over_ranks, stream)
    prove = None
    if not args.no_prove:
        cpu_msm_ms = cpu_ntt = None
        if rank == 0 and not args.no_cpu:
            # bounded CPU samples for the prove leg's component-sum baseline
            ns = 1 << 16
            pts16 = key.read(0, ns)
            t0 = time.perf_counter()
            co.msm("bn254", pts16, host_sets[0][:ns])
            cpu_msm_ms = (time.perf_counter() - t0) * 1e3 * (n / ns)
            cpu_ntt = ntt["cpu_baseline"]["value"] if ntt and ntt.get("cpu_baseline") else None
        prove = prove_leg(ctx, co, torch, args, world, rank, barrier, max_over_ranks, stream, cpu_msm_ms, cpu_ntt)
    mpc = None if args.no_mpc else mpc_leg(ctx, co, torch, args, world, rank, barrier, max_over_ranks, stream)
    stop.set()
    sampler.join(timeout=2)
    clocks = _clock_summary(rows)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel -------------------------------------------------------------
    hbm_peak, peak_src = _peaks()
    dom = max(prof_dom.items(), key=lambda kv: kv[1][1])  # measured inside the timed region
    dom_name, (dom_cnt, dom_ms) = dom
    dom_avg_ms = dom_ms / max(dom_cnt, 1)
    kernel_ms_step = sum(v[1] for v in prof.values()) / args.steps
    alg_bytes = 96.0 * n  # 32 B scalar + 64 B affine point per pair
    achieved_gbs = alg_bytes / (dom_avg_ms * 1e-3) / 1e9
    # integer roof: measured by the library's own micro-benchmarks on this GPU, right now
    imad_rate = ctx.microbench(0)
    mul_rate = ctx.microbench(1)
    W = (254 + 1 + key.window_bits - 1) // key.window_bits
    adds = float(n) * W                      # mixed additions in the accumulate kernel (upper bound: zero digits skip)
    limb_products = adds * 10 * 136          # 8M+2S per mixed add, 2N^2+N 32x32 products per 256-bit Montgomery product
    int_achieved = limb_products / (dom_avg_ms * 1e-3)
    traffic = _traffic(dom_name)

    # ---- CPU baseline leg (oracle restatement on this box's cores; bounded sample) ------------------
    cpu = None
    if not args.no_cpu:
        threads = co.max_threads()
        log_s = 18
        ns = 1 << log_s
        pts = key.read(0, ns)
        t0 = time.perf_counter()
        co.msm("bn254", pts, host_sets[0][:ns])
        dt = time.perf_counter() - t0
        cpu = {"value": dt * 1e3 * (n / ns), "unit": "ms", "cores": min(threads, co.msm_windows(ns, 254)), "kind": "port",
               "sample": "1 MSM of 2^%d pairs (same key and scalars), %d OpenMP threads, ark-ec-style window parallelism; "
                         "scaled linearly x%d to 2^20" % (log_s, threads, n // ns)}

    line = {
        "metric": METRIC, "value": ms_step, "unit": "ms", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": False, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32 limbs (8-limb Montgomery mod 254-bit p, IMAD pipe)", "data": "synthetic",
        "config": {
            "workload": "KZG commit MSM (msm_bigint), BN254 G1, 2^20 uniform canonical scalars per GPU, key = [beta^i]G "
                        "resident with precomputed window tables (c=%d, %d windows, 1 bucket set)" % (key.window_bits, W),
            "pairs_total": n * world, "sharding": "point range per rank + 1 NCCL all-gather (128 B/rank)" if world > 1 else "none",
            "l2": "inputs larger than L2: rotates %d scalar vectors (%d MB) and gathers from %d MB of tables"
                  % (N_SETS, N_SETS * 32, (W * n * 64) >> 20),
            "key_build_s": round(t_key, 3),
        },
        "pairs_per_s": n * world / (ms_step * 1e-3),
        "roofline": {"bound": "hbm", "kernel": dom_name, "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                     "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                     "note": "MSM is integer-multiply bound, not HBM bound (SURVEY 8d); see roofline_int"},
        "roofline_int": {"bound": "imad", "kernel": dom_name, "achieved": int_achieved, "peak": max(imad_rate, 136.0 * mul_rate),
                         "unit": "32x32->64 limb products/s", "frac": int_achieved / max(imad_rate, 136.0 * mul_rate),
                         "peak_source": "max of jf_microbench(0) (independent IMAD.WIDE.U32 chains, %.3e/s) and 136 x jf_microbench(1) "
                                        "(dependent Montgomery products in registers), both measured in this run" % imad_rate,
                         "mont_mul_peak_per_s": mul_rate, "mont_mul_achieved_per_s": adds * 10 / (dom_avg_ms * 1e-3),
                         "kernel_ms": dom_avg_ms, "kernel_share_of_step": dom_ms / args.steps / ms_step,
                         "kernel_share_of_kernel_time": dom_ms / args.steps / kernel_ms_step,
                         "note": "achieved = ALGORITHMIC limb products (N W mixed additions x 10 Fq products x 136, SURVEY 8d) / kernel time; "
                                 "the kernel itself issues fewer: its 2 squarings per addition take 108 wide products each"},
        "kernels_ms_per_step": {k: v[1] / args.steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])},
        "kernels_ms_note": "from a separate pass with every launch bracketed by events (that pass is ~0.2 ms slower per step "
                           "than the timed region, where only the dominant kernel is bracketed)",
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_ms, "unit": "ms", "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 128 * world},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "ntt": ntt,
        "prove": prove,
        "mpc": mpc,
    }
    _emit(line)
    if world > 1:
        dist.destroy_process_group()


def co_modulus():
    return 21888242871839275222246405745257275088548364400416034343698204186575808495617


_REAL_STDOUT = None


def _emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else any library prints (NCCL's version banner,
    make output of the oracle build, ...) has been routed to stderr."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-ntt", action="store_true", help="skip the NTT 2^22 leg (second metric)")
    ap.add_argument("--no-mpc", action="store_true", help="skip the collaborative-prover share-wise leg (config 5 shape)")
    ap.add_argument("--no-prove", action="store_true", help="skip the 2^20-gate prove leg (first metric)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
