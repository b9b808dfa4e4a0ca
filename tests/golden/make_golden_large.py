#!/usr/bin/env python3
"""Generate tests/golden/large.json: SEED-A golden proofs at the sizes bench.py reports (2^16 and 2^20 multipliers).

Same convention and same Python protocol statements as make_golden.py (oracle/bp_oracle.py); the
three O(n) primitives run through the C restatement (oracle/fast.py -> oracle/c/bp_ref.c), each of which
tests/test_c_oracle.py checks against its Python original. Nothing here touches ark_bulletproofs_b200/csrc.
Run:  python tests/golden/make_golden_large.py [name ...]      (2^20: several minutes on 8 cores)"""
import hashlib
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
sys.path.insert(0, os.path.join(HERE, ".."))
import bp_oracle as O      # noqa: E402
import fast                # noqa: E402
import oracle_cases as C   # noqa: E402

LARGE_CASES = [
    # (name, curve, kind, params)
    ("chain_2p16", "secq256k1", "chain", {"N": 1 << 16}),
    ("chain_2p16_padded", "secq256k1", "chain", {"N": 50000}),          # 15 536 padding rows
    ("shuffle_32769", "secq256k1", "shuffle", {"k": 32769, "seed": 32769}),   # 2^16 multipliers, two-phase, m = 65 538
    ("zorro_chain_2p14", "zorro", "chain", {"N": 1 << 14}),
    ("chain_2p20", "secq256k1", "chain", {"N": 1 << 20}),
]

if __name__ == "__main__":
    fast.install()
    path = os.path.join(HERE, "large.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    want = set(sys.argv[1:])
    gens_cache = {}
    for name, curve, kind, params in LARGE_CASES:
        if want and name not in want:
            continue
        cv = O.CURVES[curve]
        cap = C.gens_capacity(kind, params)
        t0 = time.time()
        key = (curve, cap)
        if key not in gens_cache:
            gens_cache.clear()
            gens_cache[key] = (O.PedersenGens(cv), fast.parallel_gens(cv, cap) if curve == "secq256k1" and cap >= 4096 else O.BulletproofGens(cv, cap, 1))
        pc, bp = gens_cache[key]
        t1 = time.time()
        proof, coms = C.oracle_prove_case(kind, params, cv, pc, bp)
        b = proof.to_bytes(cv)
        com_bytes = b"".join(O.ser_point(cv, V, True) for V in coms)
        out[name] = {
            "curve": curve, "kind": kind, "params": params, "gens_capacity": cap,
            "proof_hex": b.hex(), "sha256": hashlib.sha256(b).hexdigest(),
            "n_commitments": len(coms), "commitments_sha256": hashlib.sha256(com_bytes).hexdigest(),
        }
        print("%-18s %5d bytes  %s  (gens %.0fs, prove %.0fs)" % (name, len(b), out[name]["sha256"][:16], t1 - t0, time.time() - t1), flush=True)
        json.dump(out, open(path, "w"), indent=1)
