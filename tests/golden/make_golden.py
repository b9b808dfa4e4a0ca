#!/usr/bin/env python3
"""Generate tests/golden/proofs.json with the CPU oracle (oracle/bp_oracle.py).

The reference holds no golden bytes (all its tests use thread_rng), so these vectors are emitted
by the restatement under the SEED-A convention of SURVEY.md Appendix B: one
ChaCha20Rng::from_seed([0,1,..,31]); commitment blindings Fr::rand(rng) in commit order; the same
rng is then passed to prove(). V1/V2/V3 reproduce Appendix B's hashes.
Run:  python tests/golden/make_golden.py     (a few minutes; pure Python)"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))
import bp_oracle as O      # noqa: E402
import oracle_cases as C   # noqa: E402

out = {}
gens_cache = {}
for name, curve, kind, params in C.GOLDEN_CASES:
    cv = O.CURVES[curve]
    cap = C.gens_capacity(kind, params)
    key = (curve, cap)
    if key not in gens_cache:
        gens_cache[key] = (O.PedersenGens(cv), O.BulletproofGens(cv, cap, 1))
    pc, bp = gens_cache[key]
    t0 = time.time()
    proof, coms = C.oracle_prove_case(kind, params, cv, pc, bp)
    b = proof.to_bytes(cv)
    out[name] = {
        "curve": curve, "kind": kind, "params": params, "gens_capacity": cap,
        "proof_hex": b.hex(), "sha256": hashlib.sha256(b).hexdigest(),
        "commitments_hex": [O.ser_point(cv, V, True).hex() for V in coms],
    }
    print("%-16s %5d bytes  %s  (%.1fs)" % (name, len(b), out[name]["sha256"][:16], time.time() - t0), flush=True)
json.dump(out, open(os.path.join(HERE, "proofs.json"), "w"), indent=1)
