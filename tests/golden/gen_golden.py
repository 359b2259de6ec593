#!/usr/bin/env python3
"""tests/golden/gen_golden.py -- regenerates the committed golden fixtures.  Needs /root/reference
(parses mcl's known-answer tests out of third-parties/mcl/test/bls12_test.cpp and runs the
reference compiled by oracle/Makefile).  Outputs, all small JSON with hex strings:

  mcl_kat.json   mcl's own KATs: BLS12-381 generators + e(g1,g2) (bls12_test.cpp:19-65), the
                 finalExp input/output pair (:398-436)
  keys.json      synthetic PS keys (n = 5, 10, 20, 50) with known exponents, seed 1,
                 g = H1("abc"), gg = H2("edf"), all points normalized
  keys_bn254.json  the same for BN254 (n = 5), written when run with PSB_CURVE=bn254 (only this file is written then)
  protocol.json  reference outputs of the protocol entry points on small seeded batches:
                 verify (verdict + fused GT), randomize (serialized)
  elpasso.json   reference outputs of el_passo_provide_id (serialized credentials) and
                 el_passo_verify_id / _without_id_retrieval (verdicts) on small seeded batches
  prover.json    reference outputs of el_passo_request_id, unblind_credential and el_passo_prove_id
                 [_without_id_retrieval] under per-lane seeded streams, with the scalars those methods drew
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from tests import workload  # noqa: E402

MCL_TEST = "/root/reference/third-parties/mcl/test/bls12_test.cpp"


def hx(a):
    return np.ascontiguousarray(a).tobytes().hex()


def mcl_kat():
    src = open(MCL_TEST).read()
    blk = src[src.index("mcl::BLS12_381,"):src.index("CYBOZU_TEST_AUTO(size)")]
    nums = re.findall(r'"((?:0x[0-9A-Fa-f]+ ?)+)"', blk)
    flat = " ".join(nums).split()
    p, r = flat[0], flat[1]
    g2 = flat[2:6]
    g1 = flat[6:8]
    e = flat[8:20]
    fe = src[src.index("CYBOZU_TEST_AUTO(finalExp)"):]
    e0 = re.findall(r'"([0-9A-F]{96})', fe[fe.index("e0Str"):fe.index("e1Str")])
    e1 = re.findall(r'"([0-9A-F]{96})', fe[fe.index("e1Str"):fe.index("Fp12 e0, e1, e2;")])
    assert len(e) == 12 and len(e0) == 12 and len(e1) == 12
    return {"source": "third-parties/mcl/test/bls12_test.cpp:19-65,398-436",
            "p": p, "r": r, "g1": g1, "g2": g2, "e_g1_g2": e,
            "final_exp_in": ["0x" + x for x in e0], "final_exp_out": ["0x" + x for x in e1]}


def keys(sizes=(5, 10, 20, 50)):
    out = {}
    for n in sizes:
        k = ref.KeyMaterial(n, seed_=1)
        out[str(n)] = {"g": hx(k.g), "gg": hx(k.gg), "XX": hx(k.XX), "Y": hx(k.Y), "YY": hx(k.YY), "X": hx(k.X),
                       "x": hex(ref.fr_to_ints(k.x)[0]), "y": [hex(v) for v in ref.fr_to_ints(k.y)]}
    return {"seed": 1, "g_label": "abc", "gg_label": "edf", "layout": "mcl raw Montgomery limbs, little-endian", "keys": out}


def protocol():
    out = {}
    wl = workload.make_verify_workload(n_attrs=5, lanes=12, seed=2, tamper_every=3)
    v, gt = workload.expected_verify(wl, want_gt=True)
    out["verify"] = {"n": 5, "key_seed": 1, "sig1": hx(wl.sig1), "sig2": hx(wl.sig2),
                     "attrs": [[a.decode() for a in lane] for lane in wl.attrs],
                     "verdict": v.tolist(), "gt": hx(gt), "tampered": wl.tampered.tolist()}
    ref.seed(6)
    t = ref.fr_rand(12)
    o1, o2, ser = ref.randomize(wl.sig1, wl.sig2, t)
    out["randomize"] = {"t": hx(t), "ser": hx(ser)}
    out["hash"] = {"attr0": hx(ref.fr_set_hash_of(b"attr0")), "empty": hx(ref.fr_set_hash_of(b""))}
    return out


def elpasso():
    """reference outputs of el_passo_provide_id / el_passo_verify_id[_without_id_retrieval] (key seed 1, n = 5,
    attributes 0 and 1 hidden), incl. tampered lanes."""
    out = {}
    wl = workload.make_issuance_workload(5, 8, 2, seed=4, tamper_every=3)
    v, s1, s2, ser = workload.expected_provide_id(wl)
    ser[~v.astype(bool)] = 0
    out["provide_id"] = {"n": 5, "key_seed": 1, "A": hx(wl.A), "c": hx(wl.c), "rs": hx(wl.rs), "u": hx(wl.u),
                         "attrs": [[a.decode() for a in lane] for lane in wl.req_attrs],
                         "ads": [a.decode() for a in wl.ads], "verdict": v.tolist(), "ser": hx(ser),
                         "tampered": wl.tampered.tolist()}
    for name, with_id in (("verify_id", True), ("verify_id_without_id_retrieval", False)):
        sw = workload.make_signon_workload(5, 8, 2, seed=3, with_id=with_id, tamper_every=2)
        ev = workload.expected_verify_id(sw)
        d = {k: hx(sw.proof[k]) for k in ("sig1", "sig2", "k", "phi", "E1", "E2", "c", "rs")}
        d.update({"attrs": [[a.decode() for a in lane] for lane in sw.proof_attrs], "ads": [a.decode() for a in sw.ads],
                  "service": sw.service.decode(), "service_pt": hx(sw.service_pt), "y": hx(sw.y), "g": hx(sw.g),
                  "h": hx(sw.h), "verdict": ev.tolist(), "tampered": sw.tampered.tolist()})
        out[name] = d
    return out


def prover():
    """reference prover outputs (key seed 1, n = 5, attributes 0 and 1 hidden) + the scalars drawn, in draw order."""
    out = {}
    pw = workload.make_prover_request_workload(5, 6, 2, seed=6)
    out["request_id"] = {"n": 5, "key_seed": 1, "hidden": pw.hidden.tolist(), "attrs": [[a.decode() for a in lane] for lane in pw.attrs],
                         "ads": [a.decode() for a in pw.ads], "rnd": hx(pw.rnd), "A": hx(ref.g1_op(ref.G_NORM, pw.exp_A)),
                         "c": hx(pw.exp_c), "rs": hx(pw.exp_rs), "blind_sig1": hx(pw.blind_sig1), "blind_sig2": hx(pw.blind_sig2),
                         "unblind_sig2": hx(ref.g1_op(ref.G_NORM, pw.exp_unblind2))}
    for name, with_id in (("prove_id", True), ("prove_id_without_id_retrieval", False)):
        sw = workload.make_prover_signon_workload(5, 6, 2, seed=8, with_id=with_id)
        d = {k: hx(ref.g1_op(ref.G_NORM, sw.exp[k])) for k in ("sig1", "sig2", "phi")}
        if with_id:
            d.update({k: hx(ref.g1_op(ref.G_NORM, sw.exp[k])) for k in ("E1", "E2")})
        d.update({"k": hx(ref.g2_op(ref.G_NORM, sw.exp["k"])), "c": hx(sw.exp["c"]), "rs": hx(sw.exp["rs"]),
                  "in_sig1": hx(sw.sig1), "in_sig2": hx(sw.sig2), "rnd": hx(sw.rnd), "hidden": sw.hidden.tolist(),
                  "attrs": [[a.decode() for a in lane] for lane in sw.attrs], "ads": [a.decode() for a in sw.ads],
                  "service": sw.service.decode(), "service_pt": hx(sw.service_pt), "y": hx(sw.y), "g": hx(sw.g), "h": hx(sw.h)})
        out[name] = d
    return out


if __name__ == "__main__":
    if ref.BN254:
        with open(os.path.join(HERE, "keys_bn254.json"), "w") as f:
            json.dump(keys((5,)), f, indent=1)
        print("wrote keys_bn254.json")
        sys.exit(0)
    for name, fn in (("mcl_kat.json", mcl_kat), ("keys.json", keys), ("protocol.json", protocol), ("elpasso.json", elpasso), ("prover.json", prover)):
        with open(os.path.join(HERE, name), "w") as f:
            json.dump(fn(), f, indent=1)
        print("wrote", name, os.path.getsize(os.path.join(HERE, name)), "bytes")
