#!/usr/bin/env python3
"""bench.py -- headline benchmark: batched PS signature verification, 5 attributes, 2^20 signatures
per B200 (BASELINE.json configs[1]); one JSON line on stdout.

  python bench.py --gpus N --steps K --warmup W            # our CUDA engine (torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's own mcl CPU path

A "step" is one pass of PSVerifier::verify over the whole synthetic batch.  `value` = verifications
per second with inputs resident in HBM (psb_verify_dev, CUDA events on the launching stream, max over
ranks); `e2e` = the same through the host-buffer C-ABI call (psb_verify: H2D copies of signatures,
attribute strings and offsets + D2H of the verdict bytes inside the timed region).  Lanes shard over
ranks with no data-path collective (weak scaling: 2^20 lanes per GPU).

Synthetic data: key material from tests/golden/keys.json (seeded reference keygen with known
exponents), 2^14 distinct honest signatures built with the engine's own G1 kernels, expanded to the
full batch by per-lane randomisation (psb_randomize); every 1024-th lane tampered.  The oracle is
used only as the checker of a lane sample and for the cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_ORDERS = {"bls12_381": 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001,
            # mcl's BN254 = the Barreto-Naehrig curve with z = -(2^62 + 2^55 + 1) (mcl/include/mcl/curve_type.h), not alt_bn128
            "bn254": (lambda z: 36 * z ** 4 + 36 * z ** 3 + 18 * z ** 2 + 6 * z + 1)(-((1 << 62) + (1 << 55) + 1))}
R_ORDER = R_ORDERS["bls12_381"]   # main() switches both for --curve bn254
FPW = 6                           # u64 words of an Fp (BN254: 4)
N_ATTRS = 5
# algorithmic work per verification in 32x32->64 multiply-accumulates:
FPMUL_MAC32 = 300
# SURVEY 8d's reference-algorithm figures (Fp2 product = 3 FpMul-eq, Fp2 square / Fp2 x Fp = 2) less what this engine's
# ALGORITHM no longer does -- the roofline numerator counts the work of the algorithm actually run, not the reference's:
#   Miller: 68 fixed-argument lines scaled to constant term 1 (csrc/pairing.cuh): 9 Fp2 products + 1 extra Fp2 x Fp instead of 13
#           -> 68 * (4*3 - 2) = 680 fewer
#   final exponentiation: Fermat inversion (476) -> Bernstein-Yang divsteps (30 * 78 wide MACs = 8 FpMul-eq), and per pow_z
#           63 compressed squarings save 63*3 Fp2 squarings (378) against 6 decompressions + Montgomery's trick + one Fp2
#           inversion (18 sqr + 33 mul + 14 = 149): 5 * 229 = 1145 fewer
#   Miller: the first iteration assigns the tangent line to f instead of multiplying 1 by it: 13 Fp2 products = 39 fewer
A_MILLER2 = 7673 - 680 - 39  # two-pairing Miller loop, FpMul-eq
#           hard part through (z-1)^2 (z+p) (z^2+p^2-1) + 3 (same exponent as mcl's): 7 Fp12 products, 1 cyclotomic squaring,
#           2 Frobenius maps instead of 12, 2, 3 -> 5*54 + 18 + 15 = 303 fewer
#           pow_z finishes the four clustered top bits of |z| on the full element (6 plain instead of compressed squarings:
#           +36) instead of three more decompressions (-72): 5 * 36 = 180 fewer
#           and raises to their value 105 = (2^3-1)(2^4-1) with 7 squarings + 2 products instead of 6 + 3: 5 * 36 = 180 fewer
A_FINALEXP = 6100 + 480 - (476 - 8) - 1145 - 303 - 180 - 180
A_MSM_PER_ADD = 29    # one mixed Jacobian+affine G2 addition (7 Fp2 products + 4 Fp2 squarings)
A_MSM_PER_PAIR = 17   # one affine pair sum inside a batch: 5 Fp2 products + 1 Fp2 squaring (csrc/curve.cuh, AffBatch)
A_MSM_PER_INV = 14    # the Fp2 inversion a batch shares: norm (2) + Bernstein-Yang divsteps (8) + R^3 product (1) + 2 products (+1)
# EXECUTED wide MACs per lane, counted (not derived): tests/hostsim builds the engine's own headers for the host with call
# counters on fp_mul / fp_sqr / fp_dot2 / fp_inv, and on the device each of those calls is exactly one cios::mul (2 N^2 + N
# = 300 wide MACs), cios::sqr (234), cios::dot2 (3 N^2 + N = 444) or one divstep inversion (30 batches x 78).  Pinned by
# tests/test_hostsim.py::test_executed_mac_counts.  (mul, sqr, dot2, inv) per lane, BLS12-381:
EXEC_MAC32 = (300, 234, 444, 2340)
EXEC_OPS_MILLER = (1322, 0, 4996, 0)
EXEC_OPS_FINAL = (4182, 0, 1296, 6)
MSM_AFFINE_LEVELS = 2   # psb_api.cu default (PSB_MSM_AFFINE): table entries summed pairwise in affine coordinates, twice
K_AFF_G, K_AFF_L2, K_AFF_MIN_PAIRS = 64, 64, 10   # csrc/curve.cuh


def msm_ops(n_attrs: int, window_bits: int, levels: int = MSM_AFFINE_LEVELS):
    """Work of K = XX + sum m_i YY_i for one lane with no zero digit, mirroring aff_push_fixed_mul / aff_flush / aff_flush_l2
    (csrc/curve.cuh): returns (FpMul-eq in SURVEY 8d's accounting, (mul, sqr, dot2, inv) executed Fp-level calls)."""
    nwin = (256 + window_bits - 1) // window_bits
    E = n_attrs * nwin
    st = {"pairs": 0, "edge": 0, "inv": 0, "madd": 0}

    def batch(np_):            # one batch of np_ pairs: prefix products, one inversion, np_ pair sums
        st["pairs"] += np_
        st["edge"] += 1        # the first pair of a batch has no prefix product and no 1/d product
        st["inv"] += 1

    l2 = {"cnt": 0}

    def flush_l2():
        np_ = l2["cnt"] // 2
        if np_ >= K_AFF_MIN_PAIRS:
            batch(np_)
            st["madd"] += np_
        else:
            st["madd"] += 2 * np_
        st["madd"] += l2["cnt"] & 1
        l2["cnt"] = 0

    def flush_l1(cnt, use_l2):
        np_ = cnt // 2
        if np_ >= K_AFF_MIN_PAIRS:
            batch(np_)
            if use_l2:
                for _ in range(np_):
                    l2["cnt"] += 1
                    if l2["cnt"] == K_AFF_L2:
                        flush_l2()
            else:
                st["madd"] += np_
        else:
            st["madd"] += 2 * np_
        st["madd"] += cnt & 1

    if levels == 0:
        st["madd"] = E
    else:
        pairs = E // 2
        nb = max(1, -(-pairs // K_AFF_G))
        cap = 2 * (-(-pairs // nb)) if pairs else 2
        use_l2 = levels > 1 and E >= 4 * K_AFF_MIN_PAIRS
        cnt = 0
        for _ in range(E):
            cnt += 1
            if cnt == cap:
                flush_l1(cnt, use_l2)
                cnt = 0
        flush_l1(cnt, use_l2)
        if use_l2:
            flush_l2()
    fpmul = st["madd"] * A_MSM_PER_ADD + st["pairs"] * A_MSM_PER_PAIR + st["inv"] * A_MSM_PER_INV
    # executed calls: madd = 7 Fp2 products (2 dot2 each) + 4 Fp2 squarings (2 mul each); a pair = 5 Fp2 products + 1 squaring,
    # less the 2 products the first pair of a batch skips; the shared Fp2 inversion = 1 dot2 + 1 inv (+ its R^3 product) + 2 mul
    mul = st["madd"] * 8 + st["pairs"] * 2 + st["inv"] * 3
    dot2 = st["madd"] * 14 + st["pairs"] * 10 - st["edge"] * 4 + st["inv"]
    return fpmul, (mul, 0, dot2, st["inv"])


def hbm_line(alg_bytes_per_lane: float, lanes_per_s: float, dom_traffic, dom_ms: float) -> dict:
    """the HBM side of the roofline: algorithmic bytes against the driver-measured copy bandwidth (MEASURED_PEAKS.json, else
    the profiling recipe's fallback), and the DRAM rate of the dominant kernel from the committed ncu capture"""
    peak, src = 6457.0, "B200_PROFILING.md fallback (measured copy bandwidth of this pool)"
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peak, src = float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        pass
    out = {"algorithmic_bytes_per_lane": alg_bytes_per_lane, "achieved_GBps": alg_bytes_per_lane * lanes_per_s / 1e9,
           "peak_GBps": peak, "peak_source": src, "frac": alg_bytes_per_lane * lanes_per_s / 1e9 / peak}
    if dom_traffic:
        out["dominant_kernel_dram_GBps_ncu"] = dom_traffic / (dom_ms * 1e-3) / 1e9   # thread-local spill traffic, not algorithmic
        out["dominant_kernel_dram_frac_of_peak"] = out["dominant_kernel_dram_GBps_ncu"] / peak
    return out


def exec_mac32(ops) -> int:
    return int(sum(c * m for c, m in zip(ops, EXEC_MAC32)))
TRAFFIC_FILE = "r2w_traffic.json"   # latest committed ncu DRAM-traffic capture (profiles/)


def a_verify_fpmul(n_attrs: int, window_bits: int) -> float:
    nwin = (256 + window_bits - 1) // window_bits
    return A_MILLER2 + A_FINALEXP + msm_ops(n_attrs, window_bits)[0]


def fr_hash(msg: bytes) -> int:
    """Fr::setHashOf rule (host-side scalar bookkeeping of the workload generator)."""
    bits = R_ORDER.bit_length()    # mcl setArrayMask: keep bitSize bits, one fewer if that is still >= r
    x = int.from_bytes(hashlib.sha256(msg).digest(), "little") & ((1 << bits) - 1)
    if x >= R_ORDER:
        x &= (1 << (bits - 1)) - 1
    return x


def fr_mont(vals) -> np.ndarray:
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i] = np.frombuffer(((v << 256) % R_ORDER).to_bytes(32, "little"), dtype=np.uint64)
    return out


def load_key(n: int):
    with open(os.path.join(ROOT, "tests", "golden", "keys.json" if FPW == 6 else "keys_bn254.json")) as f:
        k = json.load(f)["keys"][str(n)]
    arr = lambda h, w: np.frombuffer(bytes.fromhex(h), dtype=np.uint64).reshape(-1, w).copy()  # noqa: E731
    g1, g2 = 3 * FPW, 6 * FPW
    return dict(g=arr(k["g"], g1), gg=arr(k["gg"], g2), XX=arr(k["XX"], g2), Y=arr(k["Y"], g1), YY=arr(k["YY"], g2),
                X=arr(k["X"], g1), x=int(k["x"], 16), y=[int(v, 16) for v in k["y"]])


def fr_val(raw: np.ndarray) -> int:
    """value of an Fr given as raw Montgomery limbs (4 x u64)."""
    return int.from_bytes(raw.tobytes(), "little") * pow(1 << 256, -1, R_ORDER) % R_ORDER


def tile_packed(blob: np.ndarray, off: np.ndarray, reps: int):
    """(blob, off) of k strings -> the same k strings repeated `reps` times (numpy only, no Python loop)."""
    total, k = int(off[-1]), off.shape[0] - 1
    if reps == 1:
        return blob, off
    b = np.concatenate([np.tile(blob[:total], reps), np.zeros(8, dtype=np.uint8)])
    o = np.empty(k * reps + 1, dtype=np.uint64)
    o[:-1] = (np.tile(off[:-1], reps).reshape(reps, k) + (np.arange(reps, dtype=np.uint64) * np.uint64(total))[:, None]).reshape(-1)
    o[-1] = total * reps
    return b, o


def lane_strings(blob, off, j: int, n_attrs: int):
    """the attribute strings of lane j of a packed batch (checker side: the reference takes lists)."""
    return [bytes(blob[int(off[j * n_attrs + i]):int(off[j * n_attrs + i + 1])]) for i in range(n_attrs)]


def attr_lists(n_attrs: int, lanes: int):
    return [[b"a%d:%d" % (i, j) for i in range(n_attrs)] for j in range(lanes)]


def honest_credentials(pkg, key, n_attrs: int, lanes: int, seed: int, base: int):
    """`lanes` distinct honest credentials under `key`: `base` signatures built from known exponents with the engine's
    own G1 kernels (sigma1 = u g, sigma2 = s sigma1, s = x + sum y_i m_i), expanded by per-lane randomisation
    (psb_randomize).  Lane j carries the attribute list of base lane j % base.  Returns sig1, sig2 and k with sigma2 = k g
    resolved lazily (the sigma2 += g tamper needs it)."""
    base = min(base, lanes)
    rng = np.random.default_rng(seed)
    attrs = attr_lists(n_attrs, base)
    s = [(key["x"] + sum(key["y"][i] * fr_hash(a[i]) for i in range(n_attrs))) % R_ORDER for a in attrs]
    u = [int.from_bytes(rng.bytes(32), "little") % R_ORDER for _ in range(base)]
    b1 = pkg.g1_mul(key["g"], fr_mont(u))                                  # sigma1 = u g
    b2 = pkg.g1_mul(key["g"], fr_mont([a * b % R_ORDER for a, b in zip(u, s)]))  # sigma2 = s sigma1
    reps = (lanes + base - 1) // base
    t = None
    if reps == 1:
        sig1, sig2 = b1[:lanes].copy(), b2[:lanes].copy()
    else:  # make every lane distinct: (t sigma1, t sigma2) with a per-lane t
        sig1 = np.empty((lanes, b1.shape[1]), dtype=np.uint64)
        sig2 = np.empty_like(sig1)
        t = np.frombuffer(rng.bytes(32 * lanes), dtype=np.uint64).reshape(lanes, 4).copy()
        t[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)  # < r; raw limbs are simply interpreted as Montgomery form
        step = base * max(1, (1 << 21) // base)
        for a in range(0, lanes, step):
            e = min(lanes, a + step)
            r = (e - a + base - 1) // base
            sig1[a:e], sig2[a:e] = pkg.PSRequester.randomize_credential(np.tile(b1, (r, 1))[:e - a], np.tile(b2, (r, 1))[:e - a], t[a:e])

    def sigma2_exponent(j: int) -> int:
        k = u[j % base] * s[j % base] % R_ORDER
        return k * fr_val(t[j]) % R_ORDER if t is not None else k
    return sig1, sig2, attrs, sigma2_exponent


def make_batch(pkg, key, lanes: int, rank: int, base: int = 1 << 14, n_attrs: int = N_ATTRS, tamper_every: int = 1024,
               attr_period: int = 0):
    """(sig1, sig2, blob, off, expected verdict) for `lanes` lanes; every `tamper_every`-th lane is tampered, cycling
    through SURVEY 8d's four kinds: sigma1 <-> sigma2, sigma1 = 0, an attribute byte changed, sigma2 += g.
    attr_period (a multiple of base and of 4 * tamper_every that divides lanes): blob / off describe only that many lanes --
    lane j carries the strings of lane j % attr_period (a large batch is verified in chunks that share them)."""
    base = min(base, lanes)
    period = attr_period or lanes
    assert lanes % period == 0 and (period == lanes or (period % base == 0 and period % (4 * tamper_every) == 0))
    sig1, sig2, attrs, s2exp = honest_credentials(pkg, key, n_attrs, lanes, 1000 + rank, base)
    bblob, boff = pkg.pack_attrs(attrs)
    reps = (period + base - 1) // base
    blob, off = tile_packed(bblob, boff, reps)
    if reps * base != period:
        off = off[:period * n_attrs + 1].copy()
    blob = blob.copy()
    expected = np.ones(lanes, dtype=np.uint8)
    plus_g = []
    for n_t, j in enumerate(range(tamper_every - 1, lanes, tamper_every)):
        kind = n_t % 4
        if kind == 0:
            sig1[j], sig2[j] = sig2[j].copy(), sig1[j].copy()
        elif kind == 1:
            sig1[j] = 0
        elif kind == 2:
            blob[int(off[(j % period) * n_attrs])] = ord("b")   # b"a0:j" -> b"b0:j"
        else:                                               # sigma2 + g = (k + 1) g: a valid point, full pairing, must fail
            plus_g.append((j, (s2exp(j) + 1) % R_ORDER))
        expected[j] = 0
    if plus_g:
        sig2[[j for j, _ in plus_g]] = pkg.g1_mul(key["g"], fr_mont([k for _, k in plus_g]))
    return sig1, sig2, blob, off, expected


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "power_w_max": max(float(r[2]) for r in self.rows if len(r) >= 7), "samples": len(sm)}


def packed_prefix(blob, off, strings: int):
    """(blob, off) of the first `strings` strings of a packed pair (views; the reference harness only reads)."""
    return blob, np.ascontiguousarray(off[:strings + 1])


def cpu_reference_verify(key_n: int, sig1, sig2, blob, off, budget_s: float, pairings: bool = True):
    """the reference's own PSVerifier::verify (mcl) on all host threads over a bounded lane sample."""
    from oracle import ref
    km = ref.KeyMaterial(key_n, seed_=1)
    threads = ref.hw_threads()
    lanes = sig1.shape[0]
    n = min(lanes, 64 * threads)
    v, t = ref.ps_verify_packed(km, sig1[:n].copy(), sig2[:n].copy(), *packed_prefix(blob, off, n * key_n), nthreads=threads, timed=True)
    rate = n / t
    n2 = int(min(lanes, max(n, rate * budget_s)))
    if n2 > n:
        v, t = ref.ps_verify_packed(km, sig1[:n2].copy(), sig2[:n2].copy(), *packed_prefix(blob, off, n2 * key_n), nthreads=threads, timed=True)
        n = n2
    out = dict(verdict=v, lanes=n, seconds=t, threads=threads, rate=n / t, jit=ref.jit_enabled())
    if pairings:
        pair_iters = 200
        tp = ref.time_pairing(pair_iters, threads)
        out["pairings_per_s"] = pair_iters * threads / tp
    return out


def max_over_ranks(values, world: int, device=None):
    """MAX over ranks of a list of floats (device times): the whole job is as slow as its slowest GPU.
    Works on any torch.distributed backend (NCCL on the GPU box, gloo in the CPU tests)."""
    if world <= 1:
        return [float(v) for v in values]
    import torch
    import torch.distributed as dist
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def job_throughput(lanes_per_rank: int, world: int, steps: int, ms_max: float) -> float:
    """whole-job units per second: every rank processed lanes_per_rank * steps lanes in ms_max (weak scaling)."""
    return world * lanes_per_rank * steps / (ms_max * 1e-3)


# ---- the other BASELINE.json configs (configs[2..4]) -- one entry each in the line's "configs" object ----------------
# Inputs are made with the engine's OWN prover side (psb_prove_id / psb_request_id, parity-tested in tests/test_gpu_prover.py)
# from honest credentials with known exponents; the reference (oracle/_ref) only checks a lane sample and is timed on it.
def rand_fr(rng, shape) -> np.ndarray:
    t = np.frombuffer(rng.bytes(32 * int(np.prod(shape))), dtype=np.uint64).reshape(*shape, 4).copy()
    t[..., 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    return t


class Ranks:
    """what a config needs from the launch: rank, world size, a barrier and the MAX over ranks of its times.  The configs
    synchronise over a gloo side group with a timeout, so a config that throws on one rank cannot leave the others waiting
    forever inside an NCCL collective: they time out, every later config is reported as an error, the headline line survives."""

    def __init__(self, rank, world, dev, torch=None, group=None):
        self.rank, self.world, self.dev, self.torch, self.group, self.failed = rank, world, dev, torch, group, False

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier(group=self.group)
        if self.torch is not None:
            self.torch.cuda.synchronize(self.dev)

    def timed(self, fn, steps):
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            out = fn()
        return out, time.perf_counter() - t0

    def max_ms(self, values):
        if self.world <= 1:
            return [float(v) for v in values]
        import torch
        import torch.distributed as dist
        t = torch.tensor(values, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return [float(v) for v in t]


def _ref_sample_lanes(lanes: int) -> int:
    return min(lanes, 1024)


def cfg_signon(pkg, rk: Ranks, steps: int, window_bits: int, lanes: int, cpu: bool, wire: bool):
    """configs[2]: EL PASSO relying-party sign-on verification -- PSVerifier::el_passo_verify_id (src/ps-verifier.cc:37-138),
    10 attributes (2 hidden), `lanes` requests per GPU; plus the same requests arriving as IdProof::toBufferString bytes."""
    n, nh, base = 10, 2, min(1 << 12, lanes)
    key = load_key(n)
    pk = pkg.PSPubKey(key["g"], key["gg"], key["XX"], key["Y"], key["YY"], window_bits=window_bits)
    sig1, sig2, attrs, _ = honest_credentials(pkg, key, n, lanes, 3000 + rk.rank, base)
    hidden = np.zeros(n, dtype=np.uint8)
    hidden[:nh] = 1
    reps = lanes // base
    shown = [[b"" if hidden[i] else a[i] for i in range(n)] for a in attrs]
    ads = [b"sess%d" % j for j in range(base)]
    all_p, shown_p, ads_p = (tile_packed(*pkg.pack_attrs(attrs), reps), tile_packed(*pkg.pack_attrs(shown), reps),
                             tile_packed(*pkg.pack_strings(ads), reps))
    pts, ok = pkg.hash_and_map_to_g1([b"rp.example", b"ghi", b"abc", b"jkl"])
    assert ok.all()
    service_pt, y, g, h = pts[0:1], pts[1:2], pts[2:3], pts[3:4]
    rng = np.random.default_rng(3100 + rk.rank)
    rq = pkg.PSRequester(pk)
    proof = rq.el_passo_prove_id(sig1, sig2, all_p, hidden, ads_p, service_pt, y, g, h, rnd=rand_fr(rng, (lanes, nh + 5)), with_id=True)
    expected = np.ones(lanes, dtype=np.uint8)
    for t, j in enumerate(range(63, lanes, 64)):      # challenge, response, sigma2 and k of every 64th proof tampered
        kind = t % 4
        if kind == 0:
            proof["c"][j, 0] ^= np.uint64(1)
        elif kind == 1:
            proof["rs"][j, 0, 0] ^= np.uint64(1)
        elif kind == 2:
            proof["sig2"][j] = proof["sig2"][j - 1]
        else:
            proof["k"][j] = proof["k"][j - 1]
        expected[j] = 0
    ver = pkg.PSVerifier(pk)
    # host side of the timed calls: page-locked input arrays and a page-locked, reused verdict buffer (psb_host_alloc)
    proof = {k: pkg.pinned_copy(v) for k, v in proof.items()}
    shown_p, ads_p = tuple(pkg.pinned_copy(a) for a in shown_p), tuple(pkg.pinned_copy(a) for a in ads_p)
    h_verdict = pkg.pinned_empty((lanes,), np.uint8)
    call = lambda: ver.el_passo_verify_id(proof, shown_p, ads_p, service_pt, y, g, h, with_id=True, out=h_verdict)  # noqa: E731
    call()                                            # tables + staging buffers
    l0 = pkg.launch_count()
    got, dt = rk.timed(call, steps)
    launches = pkg.launch_count() - l0
    ok_local = bool(np.array_equal(got, expected))
    h2d = int(sum(v.nbytes for v in proof.values()) + shown_p[0].nbytes + shown_p[1].nbytes + ads_p[0].nbytes + ads_p[1].nbytes)
    out = {"call": "PSVerifier::el_passo_verify_id (psb_verify_id)", "n_attrs": n, "hidden": nh, "lanes_per_gpu": lanes,
           "metric": "signon_verifications_per_sec", "unit": "verifications/s", "steps": steps, "window_bits": window_bits,
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": lanes, "gpu_launches": int(launches), "tampered_lanes": int((expected == 0).sum()),
           "inputs": "proofs made by psb_prove_id from honest credentials, every lane distinct; page-locked host buffers"}
    local = [dt * 1e3, 0.0 if ok_local else 1.0]
    if wire:
        ser = tuple(pkg.pinned_copy(a) for a in pkg.idproof_serialize(proof, shown_p, pk.n))
        w_out = (pkg.pinned_empty((lanes,), np.uint8), pkg.pinned_empty((lanes,), np.uint8))
        wcall = lambda: ver.el_passo_verify_id_wire(ser, ads_p, service_pt, y, g, h, with_id=True, out=w_out)  # noqa: E731
        wcall()
        (wgot, wparsed), wdt = rk.timed(wcall, steps)
        local += [wdt * 1e3, 0.0 if (np.array_equal(wgot, expected) and wparsed.all()) else 1.0]
        out["wire"] = {"call": "psb_verify_id_ser: IdProof::toBufferString bytes parsed and decompressed on the device",
                       "h2d_bytes_per_step": int(ser[0].nbytes + ser[1].nbytes + ads_p[0].nbytes + ads_p[1].nbytes),
                       "bytes_per_proof": int(ser[1][1] - ser[1][0])}
    red = rk.max_ms(local)
    out["e2e_value"] = rk.world * lanes * steps / (red[0] * 1e-3)
    out["verdicts_ok"] = red[1] == 0.0
    if wire:
        out["wire"]["e2e_value"] = rk.world * lanes * steps / (red[2] * 1e-3)
        out["wire"]["verdicts_ok"] = red[3] == 0.0
    if cpu and rk.rank == 0:
        from oracle import ref
        km, threads, S = ref.KeyMaterial(n, seed_=1), ref.hw_threads(), _ref_sample_lanes(lanes)
        t0 = time.perf_counter()
        ev = ref.verify_id(km, {k: np.ascontiguousarray(v[:S]) for k, v in proof.items()}, [shown[j % base] for j in range(S)],
                           [ads[j % base] for j in range(S)], b"rp.example", y, g, h, True, threads)
        cpu_s = time.perf_counter() - t0
        out["cpu_reference"] = {"value": S / cpu_s, "unit": "verifications/s", "cores": threads, "kind": "reference",
                                "sample": f"first {S} lanes, PSVerifier::el_passo_verify_id via mcl"}
        out["parity"] = bool(np.array_equal(ev, got[:S]))
    pk.close()
    return out


def cfg_issuance(pkg, rk: Ranks, steps: int, window_bits: int, lanes: int, cpu: bool):
    """configs[3]: EL PASSO blind issuance -- PSSigner::el_passo_provide_id (src/ps-signer.cc:63-146), 20 attributes (2 hidden)
    -- followed by PSRequester::randomize_credential (src/ps-requester.cc:139-148) of the issued credentials."""
    n, nh, base = 20, 2, min(1 << 12, lanes)
    key = load_key(n)
    pk = pkg.PSPubKey(key["g"], key["gg"], key["XX"], key["Y"], key["YY"], X_secret=key["X"], window_bits=window_bits)
    attrs = attr_lists(n, base)
    hidden = np.zeros(n, dtype=np.uint8)
    hidden[:nh] = 1
    reps = lanes // base
    shown = [[b"" if hidden[i] else a[i] for i in range(n)] for a in attrs]
    ads = [b"sess%d" % j for j in range(base)]
    all_p, shown_p, ads_p = (tile_packed(*pkg.pack_attrs(attrs), reps), tile_packed(*pkg.pack_attrs(shown), reps),
                             tile_packed(*pkg.pack_strings(ads), reps))
    rng = np.random.default_rng(4100 + rk.rank)
    A, c, rs = pkg.PSRequester(pk).el_passo_request_id(all_p, hidden, ads_p, rand_fr(rng, (lanes, nh + 2)))
    expected = np.ones(lanes, dtype=np.uint8)
    for t, j in enumerate(range(63, lanes, 64)):
        kind = t % 3
        if kind == 0:
            c[j, 0] ^= np.uint64(1)
        elif kind == 1:
            rs[j, -1, 0] ^= np.uint64(1)
        else:
            A[j] = A[j - 1]
        expected[j] = 0
    u, t_r = rand_fr(rng, (lanes,)), rand_fr(rng, (lanes,))
    sg = pkg.PSSigner(pk)
    # host side of the timed calls: page-locked input arrays and page-locked, reused output buffers (psb_host_alloc)
    A, c, rs, u, t_r = (pkg.pinned_copy(a) for a in (A, c, rs, u, t_r))
    shown_p, ads_p = tuple(pkg.pinned_copy(a) for a in shown_p), tuple(pkg.pinned_copy(a) for a in ads_p)
    G1W = A.shape[1]
    SER = 16 * G1W // 3          # sigma1 || sigma2 compressed: 2 x 8 x (u64 words of an Fp)
    o_issue = (pkg.pinned_empty((lanes,), np.uint8), pkg.pinned_empty((lanes, G1W), np.uint64), pkg.pinned_empty((lanes, G1W), np.uint64), pkg.pinned_empty((lanes, SER), np.uint8))
    o_rnd = (pkg.pinned_empty((lanes, G1W), np.uint64), pkg.pinned_empty((lanes, G1W), np.uint64), pkg.pinned_empty((lanes, SER), np.uint8))
    issue = lambda: sg.el_passo_provide_id(A, c, rs, shown_p, ads_p, u, out=o_issue)  # noqa: E731
    issue()
    l0 = pkg.launch_count()
    (v, s1, s2, ser), dt_issue = rk.timed(issue, steps)
    rnd = lambda: pkg.PSRequester.randomize_credential(s1, s2, t_r, want_serialized=True, out=o_rnd)  # noqa: E731
    rnd()
    (o1, o2, oser), dt_rnd = rk.timed(rnd, steps)
    launches = pkg.launch_count() - l0
    ok_local = bool(np.array_equal(v, expected))
    red = rk.max_ms([dt_issue * 1e3, dt_rnd * 1e3, 0.0 if ok_local else 1.0])
    out = {"call": "PSSigner::el_passo_provide_id (psb_provide_id) then PSRequester::randomize_credential (psb_randomize)",
           "n_attrs": n, "hidden": nh, "lanes_per_gpu": lanes, "metric": "credentials_issued_per_sec", "unit": "credentials/s",
           "steps": steps, "window_bits": window_bits, "gpu_launches": int(launches), "tampered_lanes": int((expected == 0).sum()),
           "e2e_value": rk.world * lanes * steps / (red[0] * 1e-3),
           "randomize_e2e_value": rk.world * lanes * steps / (red[1] * 1e-3),
           "issue_then_randomize_e2e_value": rk.world * lanes * steps / ((red[0] + red[1]) * 1e-3),
           "h2d_bytes_per_step": int(A.nbytes + c.nbytes + rs.nbytes + u.nbytes + shown_p[0].nbytes + shown_p[1].nbytes + ads_p[0].nbytes + ads_p[1].nbytes),
           "d2h_bytes_per_step": int(v.nbytes + s1.nbytes + s2.nbytes + ser.nbytes),
           "verdicts_ok": red[2] == 0.0, "inputs": "requests made by psb_request_id, every lane distinct; page-locked host buffers"}
    if cpu and rk.rank == 0:
        from oracle import ref
        km, threads, S = ref.KeyMaterial(n, seed_=1), ref.hw_threads(), _ref_sample_lanes(lanes)
        cut = lambda a: np.ascontiguousarray(a[:S])  # noqa: E731
        t0 = time.perf_counter()
        ev, e1, e2, eser = ref.provide_id(km, cut(A), cut(c), cut(rs), [shown[j % base] for j in range(S)],
                                          [ads[j % base] for j in range(S)], cut(u), threads)
        cpu_issue = time.perf_counter() - t0
        good = ev.astype(bool)
        t0 = time.perf_counter()
        r1, r2, rser = ref.randomize(cut(s1), cut(s2), cut(t_r), nthreads=threads)
        cpu_rnd = time.perf_counter() - t0
        out["cpu_reference"] = {"value": S / cpu_issue, "randomize_value": S / cpu_rnd, "unit": "credentials/s", "cores": threads,
                                "kind": "reference", "sample": f"first {S} lanes, PSSigner::el_passo_provide_id / t*sigma via mcl"}
        out["parity"] = bool(np.array_equal(ev, v[:S]) and np.array_equal(eser[good], ser[:S][good])
                             and np.array_equal(rser[good], oser[:S][good]))
    pk.close()
    return out


def cfg_verify50(pkg, rk: Ranks, window_bits: int, lanes_total: int, cpu: bool, torch):
    """configs[4]: batched PS verification, 50 attributes, `lanes_total` signatures SHARDED over the ranks (strong scaling):
    each rank verifies lanes_total / world lanes in chunks of at most 2^21, device-resident (`value`) and from host buffers (`e2e`)."""
    n, base = 50, 1 << 12
    per = lanes_total // rk.world
    chunk = min(per, 1 << 21)
    key = load_key(n)
    t0 = time.perf_counter()
    pk = pkg.PSPubKey(key["g"], key["gg"], key["XX"], key["Y"], key["YY"], window_bits=window_bits)
    sig1, sig2, blob, off, expected = make_batch(pkg, key, per, rk.rank, base=base, n_attrs=n, attr_period=chunk)
    setup_s = time.perf_counter() - t0
    dev = rk.dev
    stream = torch.cuda.Stream(dev)    # kernels and timing events share this explicit stream
    pin = lambda a: torch.from_numpy(a).pin_memory()  # noqa: E731
    h_s1, h_s2, h_blob, h_off = pin(sig1.view(np.int64)), pin(sig2.view(np.int64)), pin(blob), pin(off.view(np.int64))
    t_s1, t_s2, t_blob, t_off = (x.to(dev) for x in (h_s1, h_s2, h_blob, h_off))
    t_verdict = torch.zeros(per, dtype=torch.uint8, device=dev)
    t_ws = torch.empty(pkg.verify_ws_bytes(pk, chunk), dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(dev)
    G1B = sig1.shape[1] * 8

    def pass_dev():
        for a in range(0, per, chunk):
            pkg.verify_dev(pk, 0, min(chunk, per - a), t_s1.data_ptr() + a * G1B, t_s2.data_ptr() + a * G1B, t_blob.data_ptr(),
                           t_off.data_ptr(), 0, t_verdict.data_ptr() + a, 0, t_ws.data_ptr(), stream.cuda_stream)

    pkg.verify_dev(pk, 0, chunk, t_s1.data_ptr(), t_s2.data_ptr(), t_blob.data_ptr(), t_off.data_ptr(), 0, t_verdict.data_ptr(), 0,
                   t_ws.data_ptr(), stream.cuda_stream)          # warm-up: one chunk
    torch.cuda.synchronize(dev)
    l0 = pkg.launch_count()
    rk.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    pass_dev()
    ev1.record(stream)
    rk.barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = pkg.launch_count() - l0
    ok_dev = bool(np.array_equal(t_verdict.cpu().numpy(), expected))
    ver = pkg.PSVerifier(pk)
    n_s1, n_s2, n_blob, n_off = h_s1.numpy().view(np.uint64), h_s2.numpy().view(np.uint64), h_blob.numpy(), h_off.numpy().view(np.uint64)
    h_verdict = torch.zeros(per, dtype=torch.uint8).pin_memory().numpy()
    ver.verify(n_s1[:chunk], n_s2[:chunk], (n_blob, n_off), out=h_verdict[:chunk])   # staging buffers at the chunk size

    def pass_host():
        for a in range(0, per, chunk):
            ver.verify(n_s1[a:a + chunk], n_s2[a:a + chunk], (n_blob, n_off), out=h_verdict[a:a + chunk])

    _, e2e_s = rk.timed(pass_host, 1)
    ok_e2e = bool(np.array_equal(h_verdict, expected))
    red = rk.max_ms([dev_ms, e2e_s * 1e3, 0.0 if (ok_dev and ok_e2e) else 1.0])
    nwin = (256 + window_bits - 1) // window_bits
    out = {"call": "PSVerifier::verify (psb_verify_dev / psb_verify)", "n_attrs": n, "lanes_total": per * rk.world,
           "lanes_per_gpu": per, "chunk_lanes": chunk, "scaling": "strong", "metric": "ps_verifications_per_sec",
           "unit": "verifications/s", "steps": 1, "window_bits": window_bits, "table_bytes": pk.table_bytes,
           "value": per * rk.world / (red[0] * 1e-3), "e2e_value": per * rk.world / (red[1] * 1e-3),
           "ms_per_pass": red[0], "gpu_launches": int(launches), "verdicts_ok": red[2] == 0.0,
           "tampered_lanes": int((expected == 0).sum()), "setup_seconds": setup_s,
           "h2d_bytes_per_step": int(sig1.nbytes + sig2.nbytes + (per // chunk) * (int(off[-1]) + off.nbytes)), "d2h_bytes_per_step": per,
           "whole_step_frac_fpmul_eq": a_verify_fpmul(n, window_bits),
           "inputs": f"every signature distinct; the attribute strings of one chunk ({chunk} lanes) are shared by all chunks"}
    if cpu and rk.rank == 0:
        cb = cpu_reference_verify(n, sig1, sig2, blob, off, 3.0, pairings=False)
        out["cpu_reference"] = {"value": cb["rate"], "unit": "verifications/s", "cores": cb["threads"], "kind": "reference",
                                "sample": f"first {cb['lanes']} lanes, PSVerifier::verify via mcl"}
        out["parity"] = bool(np.array_equal(cb["verdict"], h_verdict[:cb["lanes"]]))
    del t_s1, t_s2, t_blob, t_off, t_ws, t_verdict
    torch.cuda.empty_cache()
    pk.close()
    return out


def run_configs(pkg, args, rk: Ranks, torch):
    """configs[2..4] of BASELINE.json, each through the public batched call; a config that fails on any rank is reported
    as an error entry instead of taking the headline line down with it."""
    which = [c for c in args.configs.split(",") if c and c != "none"]
    steps = max(1, min(args.steps, 3))
    cpu = not args.no_cpu_baseline
    res = {}
    jobs = {"cfg3_signon": lambda: cfg_signon(pkg, rk, steps, args.config_window_bits, args.config_lanes, cpu, True),
            "cfg4_issuance": lambda: cfg_issuance(pkg, rk, steps, args.config_window_bits, args.config_lanes, cpu),
            "cfg5_verify50": lambda: cfg_verify50(pkg, rk, args.window_bits50, args.lanes50, cpu, torch)}
    for name, job in jobs.items():
        if name.split("_")[0] not in which:
            continue
        t0 = time.perf_counter()
        if rk.failed:
            res[name] = {"error": "skipped: an earlier config failed on some rank"}
            continue
        try:
            res[name] = job()
        except Exception as e:  # noqa: BLE001  (reported in the line; the side group's timeout frees the other ranks)
            rk.failed = True
            res[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
        res[name]["wall_seconds"] = time.perf_counter() - t0
    return res


def run_in_process(pkg, args, world, key, sig1, sig2, blob, off, expected, torch):
    """SURVEY 8e's other launch model, measured after the torchrun ranks have gone: ONE process drives all `world` GPUs
    (psb_init(devices...): one host thread + stream per device inside psb_verify, verdict bytes gathered into the caller's
    array) on world x lanes lanes from host buffers."""
    t_wait = time.perf_counter()
    for d in range(1, world):       # the other ranks release their devices when they exit
        while time.perf_counter() - t_wait < 120:
            free, total = torch.cuda.mem_get_info(d)
            if free > 0.9 * total:
                break
            time.sleep(0.5)
    pkg.shutdown()
    pkg.init(list(range(world)))
    pk = pkg.PSPubKey(key["g"], key["gg"], key["XX"], key["Y"], key["YY"], window_bits=args.window_bits)
    N = sig1.shape[0]
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
    b_blob, b_off = tile_packed(blob, off, world)
    h_s1, h_s2 = pin(np.tile(sig1, (world, 1)).view(np.int64)).view(np.uint64), pin(np.tile(sig2, (world, 1)).view(np.int64)).view(np.uint64)
    h_blob, h_off = pin(b_blob), pin(b_off.view(np.int64)).view(np.uint64)
    h_verdict = torch.zeros(N * world, dtype=torch.uint8).pin_memory().numpy()
    ver = pkg.PSVerifier(pk)
    ver.verify(h_s1, h_s2, (h_blob, h_off), out=h_verdict)
    steps = max(1, min(args.steps, 3))
    l0 = pkg.launch_count()
    t0 = time.perf_counter()
    for _ in range(steps):
        ver.verify(h_s1, h_s2, (h_blob, h_off), out=h_verdict)
    dt = time.perf_counter() - t0
    out = {"call": "psb_init(all devices) + psb_verify from host buffers in ONE process (host thread + stream per device)",
           "devices": int(pkg.lib().psb_num_devices()), "lanes": N * world, "steps": steps, "value": N * world * steps / dt,
           "unit": "verifications/s", "gpu_launches": int(pkg.launch_count() - l0),
           "verdicts_ok": bool(np.array_equal(h_verdict, np.tile(expected, world)))}
    pk.close()
    return out


def run_bn254_child(args, local_rank: int):
    """the headline shape on BN254 (what initPairing() selects in the reference's shipped tests, test/ps-tests.cc:142):
    one curve per process, so a child process runs this file with --curve bn254 on the same GPU once this one is done."""
    cmd = [sys.executable, os.path.abspath(__file__), "--curve", "bn254", "--gpus", "1", "--steps", str(max(1, min(args.steps, 3))),
           "--warmup", str(min(args.warmup, 3)), "--lanes", str(args.lanes), "--window-bits", str(args.window_bits), "--cpu-budget-s", "3"]
    if args.no_cpu_baseline:
        cmd.append("--no-cpu-baseline")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "PSB_LIB")}
    env["CUDA_VISIBLE_DEVICES"] = os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local_rank] if os.environ.get("CUDA_VISIBLE_DEVICES") else str(local_rank)
    try:
        r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
        child = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        tail = (r.stderr or "")[-600:] if "r" in locals() else ""
        return {"error": f"{type(e).__name__}: {e}"[:200], "child_stderr_tail": tail}
    keep = {k: child.get(k) for k in ("metric", "value", "unit", "n_gpus", "steps", "ms_per_step", "e2e", "pairings", "gpu_launches", "cpu_baseline")}
    keep["config"] = child.get("config")
    keep["phase_ms"] = (child.get("roofline") or {}).get("phase_ms")
    keep["parity"] = (child.get("cpu_baseline") or {}).get("verdicts_agree_with_gpu")
    return keep


def _claim_stdout():
    """stdout must carry exactly ONE JSON line: native libraries (NCCL prints its version banner on the first
    collective) write to fd 1 directly, so fd 1 is pointed at stderr for the whole run and the result line goes to
    the saved descriptor."""
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    return saved


def _emit(saved_fd: int, line: dict):
    os.write(saved_fd, (json.dumps(line) + "\n").encode())


def main():
    out_fd = _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--lanes", type=int, default=1 << 20, help="lanes per GPU (the named config is 2^20)")
    ap.add_argument("--window-bits", type=int, default=20, help="fixed-base window of the per-key G2 tables (20: 13 additions per base, 1.3 GB per base)")
    ap.add_argument("--cpu-budget-s", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--configs", default="cfg3,cfg4,cfg5,inproc,bn254",
                    help="secondary entries of the line: BASELINE configs[2..4], the in-process multi-device run (N > 1), BN254; 'none' = headline only")
    ap.add_argument("--config-lanes", type=int, default=1 << 18, help="lanes per GPU of cfg3 / cfg4 (named: 2^18)")
    ap.add_argument("--config-window-bits", type=int, default=20, help="window of the per-key tables of cfg3 / cfg4 (20: +2.8 % sign-on, +6.5 % issuance over 16, profiles/r2z_ab_config_windows.txt)")
    ap.add_argument("--lanes50", type=int, default=1 << 24, help="TOTAL lanes of cfg5 (named: 2^24), sharded over the ranks")
    ap.add_argument("--window-bits50", type=int, default=20, help="fixed-base window of the 50-attribute key (20: 65 GB of tables)")
    ap.add_argument("--curve", default="bls12_381", choices=["bls12_381", "bn254"],
                    help="bn254: the headline shape on libpsb_bn254.so (the curve the reference's shipped tests run); no roofline figures")
    args = ap.parse_args()
    global R_ORDER, FPW
    bls = args.curve == "bls12_381"
    if not bls:      # one curve per process (like mcl): select the library and the reference build before either is loaded
        os.environ["PSB_CURVE"] = "bn254"
        R_ORDER, FPW = R_ORDERS["bn254"], 4
        args.configs = "none"

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"ps_verify n_attrs={N_ATTRS} lanes_per_gpu={args.lanes} " + ("BLS12-381" if bls else "BN254")

    if args.impl == "reference":
        if rank != 0:
            return
        import __graft_entry__ as ge
        key = load_key(N_ATTRS)
        from oracle import ref
        from tests import workload as wlmod
        km = ref.KeyMaterial(N_ATTRS, seed_=1)
        threads = ref.hw_threads()
        sample = 256 * threads
        attrs = wlmod.attr_strings(N_ATTRS, sample)
        sig1, sig2 = wlmod.sign_lanes(km, attrs, seed=2, nthreads=threads)
        blob, off = ref.pack_attrs(attrs)
        times = []
        for it in range(args.warmup + args.steps):
            v, t = ref.ps_verify_packed(km, sig1, sig2, blob, off, nthreads=threads, timed=True)
            assert v.all()
            if it >= args.warmup:
                times.append(t)
        val = sample * len(times) / sum(times)
        line = {"impl": "reference", "metric": "ps_verifications_per_sec", "value": val, "unit": "verifications/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u64-limbs (mcl, Xbyak JIT=%d)" % int(ref.jit_enabled()), "data": "synthetic",
                "config": {"workload": workload},
                "reference_path": "PSVerifier::verify via mcl (oracle/_ref/libpsref.so), the unmodified reference compiled by oracle/Makefile",
                "cpu_baseline": {"value": val, "unit": "verifications/s", "cores": threads, "kind": "reference",
                                 "sample": f"{sample} lanes per step"},
                "e2e": {"value": val, "unit": "verifications/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        _emit(out_fd, line)
        return

    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge
    pkg = ge.load_package()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the engine has no CPU path")
    torch.cuda.set_device(local_rank)
    side = None
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        side = dist.new_group(backend="gloo", timeout=datetime.timedelta(seconds=900))   # see class Ranks
    pkg.init([local_rank])
    dev = torch.device("cuda", local_rank)

    key = load_key(N_ATTRS)
    pk = pkg.PSPubKey(key["g"], key["gg"], key["XX"], key["Y"], key["YY"], window_bits=args.window_bits)
    N = args.lanes
    sig1, sig2, blob, off, expected = make_batch(pkg, key, N, rank)

    # measured integer-MAC rates of this GPU (roofline denominators), a few hundred ms.  A MAC32 on this part is ONE IMAD.WIDE
    # (32 x 32 + 64 -> 64); the pipe issues one per 4 clocks and scheduler (IMAD: 2 clocks per warp on the 16-lane fma pipe, the
    # wide form two passes), i.e. 32 MAC32 / clk / SM.  Valid probes: carry-chained rows in the multiplier's own form (kind 7),
    # mad.lo.cc + madc.hi pairs (kind 4), carry-free rows of 13 distinct limbs (kind 9); `peak` is the best of them.
    # Round 1's "carry-free peak" (kind 5, 44.7 / clk / SM) is NOT a MAC rate: both factors are shared by its accumulators and
    # ptxas folds the trip into one IMAD.WIDE plus 64-bit additions (SASS checked, profiles/r2a_microbench.json); it is kept
    # only as `peak_r1_invalid` so that round 1's fractions can be compared.
    geom = 148 * 8 * 256
    rate = lambda kind, iters, per: geom * iters * per / (min(pkg.microbench(kind, 148 * 8, 256, iters) for _ in range(2)) * 1e-3)  # noqa: E731
    peak_r1 = max(rate(5, 20000, 8), rate(4, 20000, 8))
    peak_chain = rate(7, 5000, 36)
    peak_mac = max(peak_chain, rate(4, 20000, 8), rate(9, 5000, 13))
    peak_issue = 32 * 148 * 1.965e9

    # ---- device-resident arm -------------------------------------------------------------------
    t_s1 = torch.from_numpy(sig1.view(np.int64)).to(dev)
    t_s2 = torch.from_numpy(sig2.view(np.int64)).to(dev)
    t_blob = torch.from_numpy(blob).to(dev)
    t_off = torch.from_numpy(off.view(np.int64)).to(dev)
    t_verdict = torch.zeros(N, dtype=torch.uint8, device=dev)
    t_ws = torch.empty(pkg.verify_ws_bytes(pk, N), dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(dev)    # an explicit non-default stream: the kernels are launched on it, the events recorded on it
    torch.cuda.synchronize(dev)

    def step_dev():
        pkg.verify_dev(pk, 0, N, t_s1.data_ptr(), t_s2.data_ptr(), t_blob.data_ptr(), t_off.data_ptr(), 0,
                       t_verdict.data_ptr(), 0, t_ws.data_ptr(), stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        step_dev()
    torch.cuda.synchronize(dev)
    got = t_verdict.cpu().numpy()
    if not np.array_equal(got, expected):
        raise SystemExit(f"verdict mismatch on {int((got != expected).sum())} lanes")

    pkg.set_profiling(True)
    phase = np.zeros(3)
    launches0 = pkg.launch_count()
    barrier()
    with ClockSampler(local_rank) as clocks:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            step_dev()
            phase += np.array(pkg.last_phase_ms(0))  # waits for this step (steps are serial anyway)
        ev1.record(stream)
        barrier()
        ms_total = ev0.elapsed_time(ev1)
    launches = pkg.launch_count() - launches0
    pkg.set_profiling(False)
    phase /= args.steps

    # ---- end-to-end arm: host buffers through psb_verify ----------------------------------------------
    ver = pkg.PSVerifier(pk)
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731  (page-locked host buffers)
    h_s1, h_s2, h_blob, h_off = pin(sig1), pin(sig2), pin(blob), pin(off.view(np.int64)).view(np.uint64)
    h_verdict = torch.zeros(N, dtype=torch.uint8).pin_memory().numpy()
    ver.verify(h_s1, h_s2, (h_blob, h_off), out=h_verdict)  # one untimed full-size call: the library sizes its staging buffers
    e2e_steps = max(1, args.steps)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        v_e2e = ver.verify(h_s1, h_s2, (h_blob, h_off), out=h_verdict)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    if not np.array_equal(v_e2e, expected):
        raise SystemExit("e2e verdict mismatch")
    h2d = int(sig1.nbytes + sig2.nbytes + int(off[-1]) + off.nbytes)
    d2h = int(N)

    # ---- pairings/s (BASELINE.json: "pairings/sec vs mcl CPU"): batched mcl::bn::pairing through psb_pairing ------
    NP = min(N, 1 << 17)
    rng = np.random.default_rng(77 + rank)
    kk = np.frombuffer(rng.bytes(32 * NP), dtype=np.uint64).reshape(NP, 4).copy()
    kk[:, 3] &= np.uint64(0x0FFFFFFFFFFFFFFF)
    Pp = pkg.g1_mul(key["g"], kk)                       # NP distinct G1 points
    Qq = np.ascontiguousarray(np.tile(key["YY"], (NP // N_ATTRS + 1, 1))[:NP])   # G2 points of the key, cycled
    pkg.pairing(Pp, Qq)   # full-size warm-up: the library grows its device buffers on first use at a size
    barrier()
    t0 = time.perf_counter()
    gt_pair = pkg.pairing(Pp, Qq)
    pair_s = time.perf_counter() - t0
    ms_total, e2e_ms, pair_ms = max_over_ranks([ms_total, e2e_s * 1e3, pair_s * 1e3], world, dev)

    # ---- the other configs (all ranks), then rank 0 alone: in-process multi-device run, BN254 ---------------------------
    table_bytes, ws_bytes = pk.table_bytes, len(t_ws)
    pk.close()
    del t_s1, t_s2, t_blob, t_off, t_ws, t_verdict
    torch.cuda.empty_cache()
    configs = run_configs(pkg, args, Ranks(rank, world, dev, torch, side), torch) if bls else {}
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    which = [c for c in args.configs.split(",") if c]
    if world > 1 and "inproc" in which:
        try:
            configs["in_process"] = run_in_process(pkg, args, world, key, sig1, sig2, blob, off, expected, torch)
        except Exception as e:  # noqa: BLE001
            configs["in_process"] = {"error": f"{type(e).__name__}: {e}"[:400]}
    pkg.shutdown()
    if bls and "bn254" in which:
        configs["bn254_verify"] = run_bn254_child(args, local_rank)

    value = job_throughput(N, world, args.steps, ms_total)
    e2e_val = job_throughput(N, world, e2e_steps, e2e_ms)
    names = ["k_verify_msm", "k_verify_miller", "k_verify_final"]
    nwin = (256 + args.window_bits - 1) // args.window_bits
    msm_fpmul, msm_exec = msm_ops(N_ATTRS, args.window_bits)
    work = [msm_fpmul, A_MILLER2, A_FINALEXP]
    executed = [exec_mac32(msm_exec), exec_mac32(EXEC_OPS_MILLER), exec_mac32(EXEC_OPS_FINAL)]
    dom = int(np.argmax(phase))
    achieved = work[dom] * FPMUL_MAC32 * N / (phase[dom] * 1e-3)
    kernels = [{"name": names[i], "ms": float(phase[i]), "fpmul_eq_per_lane": work[i],
                "achieved_tmac32": work[i] * FPMUL_MAC32 * N / (phase[i] * 1e-3) / 1e12,
                "frac": work[i] * FPMUL_MAC32 * N / (phase[i] * 1e-3) / peak_mac,
                "frac_of_carry_chain": work[i] * FPMUL_MAC32 * N / (phase[i] * 1e-3) / peak_chain,
                # what the lane really executes (counted on the instrumented host build, see EXEC_OPS_*): the numerator above
                # credits a lazily reduced Karatsuba tower (744 wide MACs per Fp2 product), the engine runs 888
                "executed_mac32_per_lane": executed[i],
                "executed_tmac32": executed[i] * N / (phase[i] * 1e-3) / 1e12,
                "executed_frac": executed[i] * N / (phase[i] * 1e-3) / peak_mac} for i in range(3)]
    traffic, traffic_source = None, None
    try:  # DRAM bytes of that kernel: NOT measured in this run -- the committed ncu --set full capture, scaled per lane
        with open(os.path.join(ROOT, "profiles", TRAFFIC_FILE)) as f:
            tr = json.load(f)
        traffic = tr["dram_bytes_per_launch"][names[dom]] / tr["lanes"] * N
        traffic_source = (f"profiles/{TRAFFIC_FILE} ({tr.get('date', 'n/a')}, ncu --set full of one wave of {tr['lanes']} lanes at "
                          f"commit {tr.get('commit', 'n/a')}), scaled per lane to this launch -- stale by construction, not from this run")
        for k in kernels:
            k["traffic_bytes_per_lane_ncu"] = tr["dram_bytes_per_launch"][k["name"]] / tr["lanes"]
    except Exception:
        pass
    whole = a_verify_fpmul(N_ATTRS, args.window_bits) * FPMUL_MAC32 * value / world
    line = {
        "metric": "ps_verifications_per_sec", "value": value, "unit": "verifications/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32-limbs (381-bit Montgomery, int32 MAC)", "data": "synthetic",
        "config": {"workload": workload, "window_bits": args.window_bits, "table_bytes": table_bytes,
                   "l2": "inputs + phase state (%.0f MB) exceed the 126 MB L2" % ((sig1.nbytes * 2 + ws_bytes) / 1e6),
                   "parallelism": f"lanes sharded over {world} GPU(s), no collective"},
        "e2e": {"value": e2e_val, "unit": "verifications/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "steps": e2e_steps},
        "pairings": {"value": job_throughput(NP, world, 1, pair_ms), "unit": "pairings/s", "lanes_per_gpu": NP,
                     "how": "psb_pairing end to end (H2D of G1/G2 points, Miller loop + final exponentiation, D2H of 576-byte GT)"},
        "gpu_launches": int(launches),
        "clocks": clocks.summary(),
        "roofline": {"bound": "int32-mac", "kernel": names[dom], "achieved": achieved / 1e12, "peak": peak_mac / 1e12,
                     "unit": "TMAC32/s", "frac": achieved / peak_mac, "traffic": traffic, "traffic_source": traffic_source,
                     "kernels": kernels,
                     "peak_source": "measured live on all SMs: best of the valid MAC32 probes (IMAD.WIDE.U32.X carry-chain rows, mad.lo.cc + madc.hi pairs, carry-free rows of 13 limbs)",
                     "peak_pipe_issue": peak_issue / 1e12, "frac_of_pipe_issue": achieved / peak_issue,
                     "peak_pipe_issue_source": "derived: 1 IMAD.WIDE per 4 clocks and scheduler = 32 MAC32 / clk / SM x 148 SMs x 1.965 GHz",
                     "peak_r1_invalid": peak_r1 / 1e12, "frac_vs_r1_denominator": achieved / peak_r1,
                     "peak_r1_invalid_note": "round 1's denominator: a probe ptxas folds into 1 IMAD.WIDE + 8 adds per trip, not a MAC rate",
                     "peak_carry_chain": peak_chain / 1e12, "frac_of_carry_chain_peak": achieved / peak_chain,
                     "peak_carry_chain_source": "measured live (IMAD.WIDE.U32.X carry-chain rows, the multiplier's instruction form)",
                     "whole_step_frac": whole / peak_mac,
                     "phase_ms": dict(zip(names, [float(x) for x in phase])),
                     "hbm": hbm_line(h2d / N + 1 + 864 * 2, value / world, traffic, phase[dom])},
    }
    if configs:
        line["configs"] = configs
    if not bls:
        line["roofline"] = {"note": "MAC32 work figures are derived for BLS12-381 only; phase_ms given", "phase_ms": line["roofline"]["phase_ms"]}
    if world == 1 and not args.no_cpu_baseline:
        try:
            cb = cpu_reference_verify(N_ATTRS, sig1, sig2, blob, off, args.cpu_budget_s)
            agree = bool(np.array_equal(cb["verdict"], got[:cb["lanes"]]))
            from oracle import ref as _ref   # checker: GT bytes of a pairing sample against mcl::bn::pairing
            pair_ok = bool(np.array_equal(gt_pair[:64], _ref.pairing(Pp[:64], Qq[:64])))
            line["pairings"]["cpu_reference_pairings_per_s"] = cb["pairings_per_s"]
            line["pairings"]["gt_bytes_agree_with_reference"] = pair_ok
            if not pair_ok:
                raise SystemExit("GPU pairing GT bytes disagree with the reference")
            line["cpu_baseline"] = {"value": cb["rate"], "unit": "verifications/s", "cores": cb["threads"], "kind": "reference",
                                    "sample": f"first {cb['lanes']} lanes of the same batch, PSVerifier::verify via mcl "
                                              f"(JIT={int(cb['jit'])}), {cb['seconds']:.1f} s",
                                    "pairings_per_s": cb["pairings_per_s"], "verdicts_agree_with_gpu": agree}
            if not agree:
                raise SystemExit("GPU verdicts disagree with the reference on the CPU sample")
        except OSError as e:  # libpsref.so absent
            line["cpu_baseline"] = {"value": None, "unit": "verifications/s", "cores": 0, "kind": "reference",
                                    "sample": f"unavailable: {e}"}
    _emit(out_fd, line)


if __name__ == "__main__":
    main()
