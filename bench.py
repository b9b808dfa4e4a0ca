#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native ark-bulletproofs hot path.

Metric (BASELINE.json): secq256k1 variable-base MSM throughput in Mpoints/s (configs[1]:
"standalone secq256k1 MSM sweep 2^12-2^24 random points/scalars on 1 B200 vs ark-ec
VariableBaseMSM on host cores"); the workload is the 2^24-point MSM of the north star.

  python bench.py --gpus N --steps K --warmup W            # this framework
  python bench.py --impl reference ...                     # the reference's CPU algorithm
                                                           # (oracle port; the Rust crate itself
                                                           #  cannot be built in this image)
One step = one MSM over one batch of synthetic points/scalars. With N > 1 (torchrun) every
rank runs its own 2^lg_n-point shard and the 64-byte partial sums are all-gathered over NCCL
and added; value = total points / max-over-ranks time ("weak" scaling).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# NCCL's own log lines (e.g. "NCCL version ..." when NCCL_DEBUG is set by the environment) must not share stdout with the
# JSON line
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

CURVE = "secq256k1"
METRIC = "secq256k1_msm_mpoints_per_s"
UNIT = "Mpoints/s"
MODMUL_IMAD = 136        # one 256-bit modmul = 136 32x32->64 multiply-adds
MADD_MODMUL = 10         # XYZZ mixed add 8M + 2S


def imad_peak():
    """Measured integer-multiply peak of this pool's B200 (tools/microbench, committed under
    profiles/); MEASURED_PEAKS.json has only HBM and bf16 figures."""
    p = os.path.join(ROOT, "profiles", "imad_peak.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["imad_per_s"], "measured: " + d.get("how", "tools/microbench")
    return 148 * 64 * 1.965e9, "nominal 148 SM x 64 IMAD/clk x 1.965 GHz (no measurement found)"


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p))["hbm_gbs"], "MEASURED_PEAKS.json"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
            # nvidia-smi needs a few hundred ms before its first row: wait for it here, outside the timed region, so that a
            # short region (5 steps x 40 ms) still holds samples
            t_end = time.time() + 5.0
            while not self.rows and time.time() < t_end and self.proc.poll() is None:
                time.sleep(0.01)
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        if not any(t0 <= ts <= t1 + 0.06 for ts, _ in self.rows):      # region shorter than one sampling period
            t0, t1 = t0 - 0.1, t1 + 0.1
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                mx = float(f[1])
                if t0 <= ts <= t1 + 0.06:
                    sm.append(float(f[0]))
                    for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                        if v.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_run(lg_sample, steps, warmup, threads=None):
    """The reference algorithm on host cores: ark-style wNAF Pippenger (oracle/c/bp_ref.c)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np

    import c_oracle
    from ark_bulletproofs_b200 import codec
    import bp_oracle as O

    # os.cpu_count(), not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1, which made the N > 1 reference arm of
    # round 1 single-threaded; the C code passes the count in a num_threads() clause, which overrides the environment
    threads = threads or os.cpu_count() or 1
    n = 1 << lg_sample
    pts = c_oracle.synth_points(0, codec.enc_point(O.SECQ256K1.G, CURVE), n, 0)
    rng = np.random.default_rng(2)
    sc = rng.integers(0, 256, size=n * 32, dtype=np.uint8)
    sc.reshape(n, 32)[:, 31] &= 0x7F
    scb = sc.tobytes()
    ptb = bytes(pts)
    for _ in range(warmup):
        c_oracle.msm_bytes(0, ptb, scb, n, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        c_oracle.msm_bytes(0, ptb, scb, n, threads)
    dt = (time.perf_counter() - t0) / steps
    return n / dt / 1e6, dt * 1e3, threads


def cpu_r1cs_baseline(lg_n, threads=None):
    """The reference's CPU schedule for one proof of 2^lg_n multipliers, assembled from bounded samples of its two hot
    operations timed on this box's host cores with the C restatement (oracle/c/bp_ref.c): the per-element generator fold
    of inner_product_proof.rs:216-225 (one 2-point msm + into_affine per output, 2(N-1) outputs for G and H together) and
    ark's Pippenger for the commitments (5N points) and the L/R cross terms (4N points over all rounds); verification is
    one MSM of 2N + 13 + 2 lg N points, timed directly. The serial TranscriptRng and the O(N) scalar loops are left
    out, so the prove figure is a lower bound. Returned for all host threads (the analogue of feature `parallel`) and
    for one thread (the crate's default features)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np

    import c_oracle
    from ark_bulletproofs_b200 import codec
    import bp_oracle as O
    N = 1 << lg_n
    threads = threads or os.cpu_count() or 1
    g = codec.enc_point(O.SECQ256K1.G, CURVE)
    rng = np.random.default_rng(7)

    def scalars(k):
        sc = rng.integers(0, 256, size=k * 32, dtype=np.uint8)
        sc.reshape(k, 32)[:, 31] &= 0x7F
        return sc.tobytes()
    threads = threads if threads else (os.cpu_count() or 1)
    out = {}
    for label, th in (("all_cores", threads), ("one_thread", 1)):
        h = 2048 if th > 1 else 256
        pts = c_oracle.synth_points(0, g, 2 * h, 0)
        t0 = time.perf_counter()
        c_oracle.fold_points(0, pts, h, scalars(1), scalars(1), th)
        t_fold = (time.perf_counter() - t0) / h                      # seconds per folded output
        n_s = 1 << (16 if th > 1 else 15)
        ptb, scb = bytes(c_oracle.synth_points(0, g, n_s, 0)), scalars(n_s)
        t0 = time.perf_counter()
        c_oracle.msm_bytes(0, ptb, scb, n_s, th)
        t_pt = (time.perf_counter() - t0) / n_s                      # seconds per MSM point at this size
        prove = t_fold * 2 * (N - 1) + t_pt * 9 * N
        verify = t_pt * (2 * N + 13 + 2 * lg_n)
        out[label] = {"threads": th, "prove_ms_lower_bound": round(prove * 1e3, 1), "verify_ms": round(verify * 1e3, 1),
                      "fold_us_per_output": round(t_fold * 1e6, 1), "msm_us_per_point": round(t_pt * 1e6, 3)}
    out["kind"] = "port"
    out["sample"] = ("oracle/c/bp_ref.c on this box: 2048 (256) folded outputs and one 2^16 (2^15)-point MSM with all cores (one thread), "
                     "scaled to the reference's operation counts for 2^%d multipliers; TranscriptRng and scalar loops excluded" % lg_n)
    return out


def r1cs_prove_verify(ctx, lg_n):
    """Secondary metric of BASELINE.json: secq256k1 R1CS prove / verify ms on the synthetic one-phase
    chain circuit with 2^lg_n multipliers, through the C ABI; the proof is
    byte-identical to the CPU oracle's at the sizes the parity tests cover."""
    from ark_bulletproofs_b200 import codec
    from ark_bulletproofs_b200 import r1cs as R
    N = 1 << lg_n
    r = codec.MODULI[CURVE][1]
    gens = R.Gens(ctx, N)
    wit = R.ChaChaRng(bytes([3] * 32))
    x0_raw = wit.scalars_raw(CURVE, 1)
    ks_raw = wit.scalars_raw(CURVE, N)
    best, best_v = None, None
    for _ in range(3):
        rng = R.ChaChaRng(bytes(range(32)))
        p = R.Prover(ctx, gens, R.Transcript(b"ChainCircuit"))
        com, var = p.commit(codec.dec_fe(x0_raw, r), rng.scalar(CURVE))
        p.chain_circuit_raw(var, N, ks_raw, x0_raw)
        t0 = time.perf_counter()
        proof = p.prove(rng)
        t_prove = (time.perf_counter() - t0) * 1e3
        st_p = ctx.last_stage_ms()
        v = R.Verifier(ctx, R.Transcript(b"ChainCircuit"))
        vv = v.commit(com)
        v.chain_circuit_raw(vv, N, ks_raw, None)
        t0 = time.perf_counter()
        v.verify(proof, gens)
        t_verify = (time.perf_counter() - t0) * 1e3
        st_v = ctx.last_stage_ms()
        if best_v is None or t_verify < best_v[0]:
            best_v = (t_verify, {k: v_ for k, v_ in st_v.items() if v_})
        if best is None or t_prove < best["prove_ms"]:
            best = {"circuit": "one-phase public-multiplier chain, 2^%d multipliers, m=1" % lg_n, "prove_ms": round(t_prove, 2),
                    "verify_ms": round(t_verify, 2), "proof_bytes": len(proof.to_bytes()),
                    "prove_stages_ms": {k: v_ for k, v_ in st_p.items() if v_}, "verify_stages_ms": {k: v_ for k, v_ in st_v.items() if v_},
                    "note": "best of 3 each; prove includes the serial TranscriptRng (8n Keccak-f on one host core, stage 'rng') that any byte-identical prover pays"}
    best["verify_ms"], best["verify_stages_ms"] = round(best_v[0], 2), best_v[1]
    return best


def kshuffle_prove_verify(ctx, k):
    """The reference's own criterion bench (benches/r1cs_secq256k1.rs:156-250): k-shuffle proof creation (2k Pedersen
    commitments + gadget + prove) and verification (2k commits + gadget + verify), through the C ABI."""
    import random
    from ark_bulletproofs_b200 import codec
    from ark_bulletproofs_b200 import r1cs as R
    cap = 1
    while cap < 2 * (k - 1):
        cap <<= 1
    gens = R.Gens(ctx, cap)
    rnd = random.Random(k)
    inp = [rnd.randrange(1 << 64) for _ in range(k)]
    out = list(inp)
    rnd.shuffle(out)
    vals_raw = codec.enc_scalars(inp + out, CURVE)

    def transcript():
        t = R.Transcript(b"ShuffleProofTest")
        t.append_message(b"dom-sep", b"ShuffleProof")
        t.append_u64(b"k", k)
        return t
    bp_, bv = None, None
    reps = []
    for _ in range(5):
        rng = R.ChaChaRng(bytes(range(32)))
        blinds_raw = rng.scalars_raw(CURVE, 2 * k)
        t0 = time.perf_counter()
        p = R.Prover(ctx, gens, transcript())
        coms_raw, vars_ = p.commit_batch_raw(vals_raw, blinds_raw, 2 * k)
        p.shuffle_gadget_native(vars_[:k], vars_[k:])
        proof = p.prove(rng)
        tp = (time.perf_counter() - t0) * 1e3
        reps.append(round(tp, 2))
        if bp_ is None or tp < bp_:
            st_best = {a: b for a, b in ctx.last_stage_ms().items() if b}
        t0 = time.perf_counter()
        v = R.Verifier(ctx, transcript())
        vv = v.commit_batch_raw(coms_raw, 2 * k)
        v.shuffle_gadget_native(vv[:k], vv[k:])
        v.verify(proof, gens)
        tv = (time.perf_counter() - t0) * 1e3
        bp_ = tp if bp_ is None or tp < bp_ else bp_
        bv = tv if bv is None or tv < bv else bv
    return {"k": k, "multipliers": 2 * (k - 1), "prove_ms": round(bp_, 2), "verify_ms": round(bv, 2), "proof_bytes": len(proof.to_bytes()),
            "prove_stages_ms": st_best, "prove_reps_ms": reps, "note": "prove_ms includes the 2k Pedersen commitments and the gadget, like the reference's bench"}


def batch_verify_bench(ctx, local_rank, rank, world, lg_n, count, nctx, dist, torch):
    """BASELINE config 5: batch verification (verifier.rs:604-691) of `count` chain-circuit proofs of 2^lg_n multipliers,
    proofs sharded over the ranks (bp_batch_verify_partial) and, inside a rank, over `nctx` contexts on host threads so
    that one context's verifier assembly overlaps another's kernels; partial points summed, NCCL all-gather for N > 1.
    `distinct` different proofs are repeated to fill the batch (the verifier's work per instance is the same)."""
    from ark_bulletproofs_b200 import Context, codec
    from ark_bulletproofs_b200 import r1cs as R
    from ark_bulletproofs_b200.dist import allgather_sum_points
    N = 1 << lg_n
    r = codec.MODULI[CURVE][1]
    distinct = 4
    ctxs = [ctx] + [Context(CURVE, local_rank) for _ in range(nctx - 1)]
    gens_k = [R.Gens(c, N) for c in ctxs]
    made = []
    for d in range(distinct):
        wit = R.ChaChaRng(bytes([(4 + d) % 256] * 32))
        x0_raw = wit.scalars_raw(CURVE, 1)
        ks_raw = wit.scalars_raw(CURVE, N)
        rng = R.ChaChaRng(bytes(range(32)))
        p = R.Prover(ctx, gens_k[0], R.Transcript(b"ChainCircuit"))
        com, var = p.commit(codec.dec_fe(x0_raw, r), rng.scalar(CURVE))
        p.chain_circuit_raw(var, N, ks_raw, x0_raw)
        made.append((p.prove(rng), com, ks_raw))
    mine = list(range(rank, count, world))
    arng = R.ChaChaRng(bytes([5] * 32))
    alphas = [arng.scalar(CURVE) for _ in range(count)]
    shares = [mine[k::nctx] for k in range(nctx)]

    def build(c, idx):
        out = []
        for i in idx:
            proof, com, ks_raw = made[i % distinct]
            v = R.Verifier(c, R.Transcript(b"ChainCircuit"))
            vv = v.commit(com)
            v.chain_circuit_raw(vv, N, ks_raw, None)
            out.append((v, proof))
        return out
    best, best_total = None, None
    totals = {}
    for it, mode in enumerate(("warmup", "host_transcript", "device_transcript", "host_transcript", "device_transcript")):
        for c in ctxs:
            c.set_device_transcript(32 if mode == "device_transcript" else 0)
        if world > 1:
            dist.barrier()
        t00 = time.perf_counter()
        parts = [None] * nctx
        t_verify = [0.0] * nctx

        nbuild = max(1, (os.cpu_count() or 1) // nctx)

        def work(k):
            # verifier assembly (the caller's circuit construction, host only) is part of the measured pipeline; it is
            # spread over the host cores: `nbuild` threads per context build disjoint parts of the context's share
            parts_k = [None] * nbuild

            def build_part(j):
                parts_k[j] = build(ctxs[k], shares[k][j::nbuild])
            bt = [threading.Thread(target=build_part, args=(j,)) for j in range(nbuild)]
            for t in bt:
                t.start()
            for t in bt:
                t.join()
            insts, order = [], []
            for j in range(nbuild):
                insts += parts_k[j]
                order += shares[k][j::nbuild]
            t0 = time.perf_counter()
            parts[k] = R.batch_verify_partial(ctxs[k], [alphas[i] for i in order], insts, gens_k[k])
            t_verify[k] = time.perf_counter() - t0
        th = [threading.Thread(target=work, args=(k,)) for k in range(nctx)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        part = ctx.points_sum(parts) if nctx > 1 else parts[0]
        if world > 1:
            _, idn = allgather_sum_points(CURVE, codec.enc_point(part, CURVE) if part is not None else bytes(64), part is None,
                                          device=torch.device("cuda", local_rank))
            assert idn, "batch rejected"
        else:
            assert part is None, "batch rejected"
        total = time.perf_counter() - t00
        if world > 1:
            t = torch.tensor([total], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t[0])
        if mode == "warmup":
            continue
        totals[mode] = min(total, totals.get(mode, 1e9))
        totals[mode + "_verify_only"] = min(max(t_verify), totals.get(mode + "_verify_only", 1e9))
        if mode == "device_transcript" and (best_total is None or total < best_total):
            best_total, best = total, max(t_verify)
    return {"proofs": count, "multipliers": "2^%d" % lg_n, "n_gpus": world, "contexts_per_gpu": nctx, "ms": round(best_total * 1e3, 1),
            "proofs_per_s": round(count / best_total, 1), "verify_only_ms": round(best * 1e3, 1),
            "ms_with_host_transcript": round(totals["host_transcript"] * 1e3, 1),
            "verify_only_ms_with_host_transcript": round(totals["host_transcript_verify_only"] * 1e3, 1),
            "note": "wall time of building the %d verifiers (constraint systems, commitments) and bp_batch_verify_partial on every context "
                    "(verify_only_ms: the slowest context's bp_batch_verify_partial alone); IPA challenges of each context's proofs derived in one "
                    "launch by the device transcript (ms_with_host_transcript: the same with the host transcript); %d distinct proofs repeated" % (count, distinct)}


def prover_throughput_bench(local_rank, lg_n, concurrencies, golden_sha=None):
    """North star "absolute ms and proofs/s": one proof keeps the GPU idle most of the time (the serial TranscriptRng Keccak
    chain runs on one host core), so throughput comes from several provers -- one context, stream and host thread each --
    sharing the GPU: the Keccak chains of different proofs run on different cores while the GPU works on another proof's
    MSMs and folds. Every prover proves the same SEED-A chain statement; all proofs must be byte-identical."""
    import hashlib
    from ark_bulletproofs_b200 import Context, codec
    from ark_bulletproofs_b200 import r1cs as R
    N = 1 << lg_n
    r = codec.MODULI[CURVE][1]
    wit = R.ChaChaRng(bytes([3] * 32))
    x0_raw = wit.scalars_raw(CURVE, 1)
    ks_raw = wit.scalars_raw(CURVE, N)
    x0 = codec.dec_fe(x0_raw, r)
    kmax = max(concurrencies)
    ctxs = [Context(CURVE, local_rank) for _ in range(kmax)]
    gens = [R.Gens(c, N) for c in ctxs]

    def one(k):
        rng = R.ChaChaRng(bytes(range(32)))
        p = R.Prover(ctxs[k], gens[k], R.Transcript(b"ChainCircuit"))
        _, var = p.commit(x0, rng.scalar(CURVE))
        p.chain_circuit_raw(var, N, ks_raw, x0_raw)
        return hashlib.sha256(p.prove(rng).to_bytes()).hexdigest()
    ref = one(0)
    if golden_sha is not None:
        assert ref == golden_sha, "proof differs from the oracle's golden bytes"
    out = []
    for conc in concurrencies:
        per = max(6, 24 // conc)      # at least six proofs per prover: three were a burst whose rate varied 2x between runs
        digs = [None] * conc

        def work(k):
            d = set()
            for _ in range(per):
                d.add(one(k))
            digs[k] = d
        th = [threading.Thread(target=work, args=(k,)) for k in range(conc)]
        t0 = time.perf_counter()
        for t in th:
            t.start()
        for t in th:
            t.join()
        dt = time.perf_counter() - t0
        assert all(d == {ref} for d in digs), "concurrent provers disagree on the proof bytes"
        out.append({"provers": conc, "proofs": conc * per, "proofs_per_s": round(conc * per / dt, 2), "ms_per_proof": round(dt * 1e3 / (conc * per), 2)})
    return {"multipliers": "2^%d" % lg_n, "host_cores": os.cpu_count(), "runs": out, "byte_identical": True,
            "golden": golden_sha is not None,
            "note": "wall time including circuit construction; one bp_ctx (stream, scratch, generator tables) and one host thread per prover"}


def other_curve_bench(curve, lg_n):
    """BASELINE config 5, second half: zorro / curve25519 R1CS prove + verify of the chain circuit at 2^lg_n multipliers."""
    from ark_bulletproofs_b200 import Context, codec
    from ark_bulletproofs_b200 import r1cs as R
    N = 1 << lg_n
    c = Context(curve, 0)
    t0 = time.perf_counter()
    gens = R.Gens(c, N)
    gens_s = time.perf_counter() - t0
    r = codec.MODULI[curve][1]
    wit = R.ChaChaRng(bytes([3] * 32))
    x0_raw = wit.scalars_raw(curve, 1)
    ks_raw = wit.scalars_raw(curve, N)
    best_p, best_v = None, None
    for _ in range(2):
        rng = R.ChaChaRng(bytes(range(32)))
        p = R.Prover(c, gens, R.Transcript(b"ChainCircuit"))
        com, var = p.commit(codec.dec_fe(x0_raw, r), rng.scalar(curve))
        p.chain_circuit_raw(var, N, ks_raw, x0_raw)
        t0 = time.perf_counter()
        proof = p.prove(rng)
        tp = (time.perf_counter() - t0) * 1e3
        st = {k: v for k, v in c.last_stage_ms().items() if v}
        v = R.Verifier(c, R.Transcript(b"ChainCircuit"))
        vv = v.commit(com)
        v.chain_circuit_raw(vv, N, ks_raw, None)
        t0 = time.perf_counter()
        v.verify(proof, gens)
        tv = (time.perf_counter() - t0) * 1e3
        if best_p is None or tp < best_p[0]:
            best_p = (tp, st)
        best_v = tv if best_v is None or tv < best_v else best_v
    return {"curve": curve, "multipliers": "2^%d" % lg_n, "prove_ms": round(best_p[0], 2), "verify_ms": round(best_v, 2),
            "rng_ms": best_p[1].get("rng"), "ipa_ms": best_p[1].get("ipa"), "gens_s": round(gens_s, 2), "proof_bytes": len(proof.to_bytes())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--lg-n", type=int, default=24, help="log2 of the MSM size per GPU")
    ap.add_argument("--cpu-lg-n", type=int, default=0, help="log2 of the CPU sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--r1cs-lg-n", default="16,20", help="log2 multipliers of the secondary R1CS prove/verify measurements, comma separated (0 = skip)")
    ap.add_argument("--batch-verify", default="16,1024,4", help="lg multipliers, proofs, contexts per GPU of the batch-verification measurement (empty = skip)")
    ap.add_argument("--other-curves", default="zorro,curve25519", help="curves of the 2^16 prove/verify measurement at N = 1 (empty = skip)")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--prover-throughput", default="16:1,4,8,16", help="lg multipliers : concurrent provers of the proofs/s measurement at N = 1 (empty = skip)")
    ap.add_argument("--shuffle-k", type=int, default=1024, help="k of the reference's own k-shuffle bench, reported next to r1cs (0 = skip)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3) if args.impl == "b200" else max(args.warmup, 1)
    workload = "secq256k1 variable-base MSM, 2^%d points per GPU, uniform 255-bit scalars" % args.lg_n

    if args.impl == "reference":
        if rank != 0:
            return 0
        # the same workload as the b200 arm (2^lg_n points per step); every step is one full MSM on all host cores, so the
        # number of timed steps is capped to keep the run within minutes (a step takes ~10 s at 2^24 on 16 cores)
        lg = args.cpu_lg_n or args.lg_n
        steps = max(1, min(args.steps, 2 if lg >= 22 else 5))
        wu = 1 if lg < 22 else 0
        val, ms, threads = cpu_reference_run(lg, steps, wu)
        line = {
            "impl": "reference", "metric": METRIC, "value": round(val, 4), "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": wu, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u256 (Montgomery, 4x64-bit limbs)", "data": "synthetic",
            "config": {"workload": workload, "steps_requested": args.steps,
                       "note": "CPU arm: same 2^%d-point MSM per step, steps capped at %d (one step is a whole MSM on every host core)" % (lg, steps)},
            "cpu_baseline": {"value": round(val, 4), "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": "oracle/c/bp_ref.c: ark-ec 0.4 msm_bigint_wnaf restated in portable C (windows over %d OpenMP threads, the "
                                       "analogue of feature `parallel`; generic CIOS multiplication, roughly 2x slower per modmul than ark-ff's "
                                       "x86-64 assembly), %d x 2^%d points; the Rust crate cannot be built here (no cargo/rustc)" % (threads, steps, lg)},
            "e2e": {"value": round(val, 4), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from ark_bulletproofs_b200 import Context

    ctx = Context(CURVE, local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda", local_rank))
    n = 1 << args.lg_n
    pts = torch.empty(n * 64, dtype=torch.uint8, device="cuda")
    ctx.synth_points_device(pts.data_ptr(), n, rank * n)
    ctx.sync()
    g = torch.Generator(device="cuda").manual_seed(2 + rank)
    sc = torch.randint(0, 256, (n * 32,), dtype=torch.uint8, device="cuda", generator=g)
    sc.view(-1, 32)[:, 31] &= 0x7F            # 255-bit raw Montgomery residues: uniform scalars < r
    torch.cuda.synchronize()
    h_pts = torch.empty(n * 64, dtype=torch.uint8, pin_memory=True)
    h_sc = torch.empty(n * 32, dtype=torch.uint8, pin_memory=True)
    h_pts.copy_(pts)
    h_sc.copy_(sc)
    torch.cuda.synchronize()
    from ark_bulletproofs_b200.dist import allgather_sum_points

    def combine(raw, ident):
        """all-gather the per-GPU partial points (NCCL), add them on every rank."""
        if world == 1:
            return raw, ident
        return allgather_sum_points(CURVE, raw, ident, device=torch.device("cuda", local_rank))

    def step_device():
        return combine(*ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n))

    def step_e2e():
        out = ctypes.create_string_buffer(64)
        idn = ctypes.c_int(0)
        ctx._check(ctx.lib.bp_msm(ctx.h, h_pts.data_ptr(), h_sc.data_ptr(), n, out, ctypes.byref(idn)))
        return combine(out.raw, bool(idn.value))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t0 = time.time()
        e0.record(stream)
        for _ in range(steps):
            res = fn()
        e1.record(stream)
        e1.synchronize()
        barrier()
        t1 = time.time()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res, t0, t1

    ctx.set_timing(True)
    for _ in range(warmup):
        ref = step_device()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    l0 = ctx.launches
    ms, res, t0, t1 = timed(step_device, args.steps)
    launches = ctx.launches - l0
    clocks = sampler.stop(t0, t1) if sampler else None
    assert res == ref, "MSM result changed between steps"
    phases = ctx.last_phases()
    ms_per_step = ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6

    step_e2e()
    ms_e2e, res_e2e, _, _ = timed(step_e2e, args.steps)
    assert res_e2e == ref, "host-buffer path disagrees with the device-resident path"
    e2e_val = world * n / (ms_e2e / args.steps * 1e-3) / 1e6

    # the same MSM with the bases resident on the GPU (bp_bases_upload; the protocol's large MSMs are over fixed generators):
    # only the 32-byte scalars cross PCIe inside the timed region
    hb = ctx.bases_upload(h_pts.data_ptr(), n)

    def step_resident():
        return combine(*ctx.msm_bases(hb, h_sc.data_ptr(), n))
    step_resident()
    ms_res, res_res, _, _ = timed(step_resident, args.steps)
    assert res_res == ref, "resident-bases path disagrees with the device-resident path"
    ctx.bases_free(hb)
    e2e_res = {"value": round(world * n / (ms_res / args.steps * 1e-3) / 1e6, 3), "unit": UNIT, "h2d_bytes_per_step": n * 32,
               "d2h_bytes_per_step": 64 + phases["windows"] * 128, "ms_per_step": round(ms_res / args.steps, 4),
               "api": "bp_msm_bases (C ABI; bases uploaded once with bp_bases_upload, scalars in pinned host memory)"}

    # BASELINE config 1: the MSM sweep 2^12 .. 2^24 on one GPU (device-resident inputs: prefixes of the same buffers)
    sweep = None
    if world == 1 and not args.no_sweep:
        sweep = []
        peak_imad = imad_peak()[0]
        for lg in range(12, args.lg_n + 1, 2):
            m = 1 << lg
            for _ in range(3):
                ctx.msm_device(pts.data_ptr(), sc.data_ptr(), m)
            reps = 20 if lg <= 18 else 5
            ms_s, _, _, _ = timed(lambda: ctx.msm_device(pts.data_ptr(), sc.data_ptr(), m), reps)
            ph = ctx.last_phases()
            per = ms_s / reps
            acc = ph["ms"].get("accumulate") or 0.0
            sweep.append({"lg_n": lg, "ms": round(per, 4), "mpoints_per_s": round(m / per / 1e3, 2), "window_bits": ph["c"], "windows": ph["windows"],
                          "accumulate_ms": round(acc, 4),
                          "accumulate_frac_of_imad_peak": round(ph["entries"] * MADD_MODMUL * MODMUL_IMAD / (acc * 1e-3) / peak_imad, 4) if acc > 0 else None,
                          "whole_msm_frac_of_imad_peak": round(ph["entries"] * MADD_MODMUL * MODMUL_IMAD / (per * 1e-3) / peak_imad, 4)})
        ctx.msm_device(pts.data_ptr(), sc.data_ptr(), n)      # leave the phase timers on the headline size
        phases = ctx.last_phases()

    # strong scaling (north star: "a 2^24-point MSM running on 1-8 B200"): the same 2^lg_n points in total, 2^lg_n / N per GPU
    strong = None
    if world > 1:
        from ark_bulletproofs_b200.dist import shard_range
        lo, hi = shard_range(n, rank, world)
        cnt = hi - lo

        def step_strong():
            return combine(*ctx.msm_device(pts.data_ptr(), sc.data_ptr(), cnt))    # this rank's points are distinct multiples of G already
        for _ in range(warmup):
            step_strong()
        ms_st, _, _, _ = timed(step_strong, args.steps)
        per = ms_st / args.steps
        strong = {"scaling": "strong", "total_points": n, "points_per_gpu": cnt, "ms_per_step": round(per, 4), "value": round(n / per / 1e3, 3), "unit": UNIT,
                  "speedup_vs_1gpu_weak_step": round(ms_per_step / per, 3),
                  "note": "same 2^%d points in total; per-GPU MSM of 2^%d/%d points + one 64-byte all-gather and a host point sum per step" % (args.lg_n, args.lg_n, world)}

    # secondary metric: R1CS prove / verify. With N > 1 the context switches to multi-GPU mode (cyclic generator
    # shards, partial points all-gathered over NCCL) and every rank takes part in the same proof.
    r1cs = None
    lgs = [int(x) for x in str(args.r1cs_lg_n).split(",") if int(x) > 0]
    if lgs:
        if world > 1:
            ctx.init_nccl(rank, world)          # library-owned ncclAllGather of the partial points (no Python on the data path)
        runs = []
        for lg in lgs:
            one = r1cs_prove_verify(ctx, lg)
            if world > 1:
                t = torch.tensor([one["prove_ms"], one["verify_ms"]], device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                one["prove_ms"], one["verify_ms"] = round(float(t[0]), 2), round(float(t[1]), 2)
            runs.append(one)
        r1cs = dict(runs[0])                     # the 2^16 circuit of BASELINE's config 2 stays at the top level
        r1cs["sizes"] = {"2^%d" % lg: {"prove_ms": o["prove_ms"], "verify_ms": o["verify_ms"], "proof_bytes": o["proof_bytes"],
                                        "rng_ms": o["prove_stages_ms"].get("rng"), "ipa_ms": o["prove_stages_ms"].get("ipa")} for lg, o in zip(lgs, runs)}
        if world > 1:
            r1cs["sharding"] = ("generators cyclic over %d GPUs; every MSM's 64 B partial points all-gathered by the library (ncclAllGather on the "
                                "context's stream, bp_ctx_init_nccl); scalars, transcript and TranscriptRng replicated" % world)
        elif args.shuffle_k > 1:
            r1cs["reference_bench_kshuffle"] = kshuffle_prove_verify(ctx, args.shuffle_k)
        if world == 1 and not args.no_cpu_baseline:
            r1cs["cpu_baseline"] = cpu_r1cs_baseline(lgs[0])

    batch = None
    if args.batch_verify:
        # batch verification shards proofs, not generators: fresh single-GPU contexts when `ctx` is in multi-GPU mode
        blg, bcount, bctx = [int(x) for x in args.batch_verify.split(",")]
        batch = batch_verify_bench(ctx if world == 1 else Context(CURVE, local_rank), local_rank, rank, world, blg, bcount, bctx,
                                   dist if world > 1 else None, torch)
    throughput = None
    if world == 1 and args.prover_throughput:
        tlg, tcs = args.prover_throughput.split(":")
        gpath = os.path.join(ROOT, "tests", "golden", "large.json")
        gold = json.load(open(gpath)).get("chain_2p%s" % tlg, {}).get("sha256") if os.path.exists(gpath) else None
        throughput = prover_throughput_bench(local_rank, int(tlg), [int(x) for x in tcs.split(",")], gold)
    others = None
    if world == 1 and args.other_curves:
        others = [other_curve_bench(c, 16) for c in args.other_curves.split(",") if c]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # roofline of the dominant kernel (msm_accumulate: integer-multiply bound)
    acc_ms = phases["ms"]["accumulate"]
    alg_imad = phases["entries"] * MADD_MODMUL * MODMUL_IMAD
    peak, peak_src = imad_peak()
    achieved = alg_imad / (acc_ms * 1e-3)
    traffic = None
    tp = os.path.join(ROOT, "profiles", "accumulate_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch_lg%d" % args.lg_n)
    hbm, hbm_src = hbm_peak()
    # pair ordering: the pipeline's bucket sort reads the 4-byte keys twice (histogram, scatter) and then moves every 8-byte
    # pair once out of the scatter and once through the bins kernel (4 + 4 + 8 + 8 + 8 = 32 B per pair); the library radix sort reads the keys once and moves every pair once per 8-bit pass
    if phases.get("bucket_sort"):
        sort_bytes = phases["entries"] * 32
        sort_kernel = "bucket sort of (bucket, point) pairs (csrc/msm_sort.cuh: hist, colscan, binscan, scatter, bins)"
    else:
        sort_bytes = phases["entries"] * (4 + 8 * 2 * ((phases["c"] - 1 + max(1, (phases["windows"] - 1).bit_length()) + 7) // 8))
        sort_kernel = "cub radix sort of (bucket, point) pairs"
    # multiplier issue slots the mixed addition really executes (cuobjdump counts, tools/microbench7.cu): a wide product
    # (IMAD.WIDE / IMAD.HI) holds the multiplier for two slots, a 32-bit IMAD for one. Product: 128 wide + 8 = 264;
    # squaring (Fp::sqr_sos): 90 wide + 10 hi + 22 = 222; fused a*b - c*d with one reduction (Fp::mul2): 184 wide + 8 hi + 8 = 392.
    # madd = 6 products + 2 squarings + 1 fused = 2420 slots (round 1: 10 products = 2640). The first addition into an
    # empty bucket is a copy.
    real_adds = phases["entries"] - phases["windows"] * (1 << (phases["c"] - 1))
    pipe_frac = real_adds * 2420 / (acc_ms * 1e-3) / peak
    line = {
        "metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u256 (Montgomery, 8x32-bit limbs, IMAD.WIDE)", "data": "synthetic",
        "config": {"workload": workload, "curve": CURVE, "window_bits": phases["c"], "windows": phases["windows"],
                   "l2": "inputs (%.2f GB) exceed the 126 MB L2; no flush needed" % (n * 96 / 1e9),
                   "parallelism": "msm-shard x%d + all-gather of 64 B partial points" % world},
        "e2e": {"value": round(e2e_val, 3), "unit": UNIT, "h2d_bytes_per_step": n * 96, "d2h_bytes_per_step": 64 + phases["windows"] * 128,
                "ms_per_step": round(ms_e2e / args.steps, 4), "api": "bp_msm (C ABI, pinned host buffers)"},
        "e2e_resident_bases": e2e_res,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "imad", "kernel": "msm_accumulate_kernel", "achieved": round(achieved / 1e12, 4), "peak": round(peak / 1e12, 4),
                     "unit": "TIMAD/s", "frac": round(achieved / peak, 4), "traffic": traffic,
                     "algorithmic": "%d bucket additions x 10 modmul x 136 IMAD" % phases["entries"], "peak_source": peak_src,
                     "launch_ms": round(acc_ms, 4),
                     # `frac` counts the reference formulation (10 modmul x 136 multiply-adds per addition) against the
                     # 32-bit IMAD peak; a 32x32->64 product holds the multiplier for two issue slots, so a formulation of
                     # 10 plain products tops out at 0.5 (round 1: 0.448). `multiplier_pipe_busy` is what the kernel really
                     # issues (2420 slots per addition after the fused product and the dedicated squaring) over the same peak.
                     "multiplier_pipe_busy": round(pipe_frac, 4)},
        "roofline_hbm": {"bound": "hbm", "kernel": sort_kernel, "achieved": round(sort_bytes / (phases["ms"]["sort"] * 1e-3) / 1e9, 1),
                         "peak": hbm, "unit": "GB/s", "frac": round(sort_bytes / (phases["ms"]["sort"] * 1e-3) / 1e9 / hbm, 4),
                         "peak_source": hbm_src, "launch_ms": round(phases["ms"]["sort"], 4)},
        "phases_ms": {k: round(v, 4) for k, v in phases["ms"].items()},
        # every MSM phase against the roofline that bounds it (algorithmic work per launch / measured phase time):
        #   digits: 32 B scalar in, 8 B (key, value) out per window          -> HBM
        #   partials / reduce: XYZZ additions of 14 modmul x 136 IMAD        -> IMAD (reduce: 2 per bucket)
        "kernels": {
            "msm_digits_kernel": {"bound": "hbm", "frac": round(n * (32 + (4 if phases.get("bucket_sort") else 8) * phases["windows"]) / (phases["ms"]["digits"] * 1e-3) / 1e9 / hbm, 4)},
            "pair_sort": {"bound": "hbm", "frac": round(sort_bytes / (phases["ms"]["sort"] * 1e-3) / 1e9 / hbm, 4)},
            "msm_accumulate_kernel": {"bound": "imad", "frac": round(achieved / peak, 4)},
            "msm_reduce_kernel+window_sum": {"bound": "imad", "frac": round(phases["windows"] * (1 << (phases["c"] - 1)) * 2 * 14 * MODMUL_IMAD / (phases["ms"]["reduce"] * 1e-3) / peak, 4)},
        },
    }
    if sweep is not None:
        line["sweep"] = sweep
    if strong is not None:
        line["strong_scaling"] = strong
    if batch is not None:
        line["batch_verify"] = batch
    if others is not None:
        line["other_curves"] = others
    if throughput is not None:
        line["prover_throughput"] = throughput
    if r1cs is not None:
        line["r1cs"] = r1cs
    if world == 1 and not args.no_cpu_baseline:
        ncores = os.cpu_count() or 1
        lg = args.cpu_lg_n or (22 if ncores >= 16 else 18)      # bounded sample; `--impl reference` runs the full 2^24 workload
        val, cms, threads = cpu_reference_run(lg, 1, 0)
        line["cpu_baseline"] = {"value": round(val, 4), "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "one 2^%d-point MSM, oracle/c/bp_ref.c (ark-ec 0.4 msm_bigint_wnaf restated in portable C, windows over "
                                          "%d OpenMP threads; roughly 2x slower per modmul than ark-ff's x86-64 assembly), %.1f s" % (lg, threads, cms / 1e3)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
