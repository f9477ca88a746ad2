#!/usr/bin/env python3
"""Mint the golden vectors under tests/golden/ with the Python oracle (oracle/pyref).

The reference holds no golden vector for this path, so these are
oracle-generated; what pins the oracle itself is listed in oracle/pyref/__init__.py and checked
by tests/test_oracle.py.  Layout = the C ABI's (include/zkb.h): Montgomery u64 limbs.

    python tests/golden/make_golden.py            # rewrites every *.npz next to this file
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle.pyref import groth16 as G  # noqa: E402
from oracle.pyref.curves import CURVES  # noqa: E402
from oracle.pyref.fields import BLS12_381, BN254, FR, splitmix64, stream_field  # noqa: E402
from oracle.pyref.msm import msm_pippenger  # noqa: E402
from oracle.pyref.ntt import Domain  # noqa: E402
from oracle.pyref.r1cs import ConstraintSystem, mimc_circuit, mini_circuit  # noqa: E402
from tests import helpers as H  # noqa: E402


def toxic(curve_id, seed):
    """alpha, beta, gamma, delta, t, r, s from SplitMix64 stream `seed`."""
    p = FR[curve_id].p
    return [stream_field(seed, i, p) for i in range(7)]


def groth16_case(name, curve_id, build, seed):
    fr = FR[curve_id]
    cs = ConstraintSystem(fr.p)
    build(cs)
    assert cs.is_satisfied()
    alpha, beta, gamma, delta, t, r, s = toxic(curve_id, seed)
    pk = G.generate_parameters(cs, curve_id, alpha, beta, gamma, delta, t)
    proof = G.create_proof(pk, cs, r, s)
    assert proof == G.proof_in_exponent(pk, cs, r, s), "oracle self-check (in-the-exponent identity) failed"
    h = G.witness_map(cs, curve_id)
    out = {"curve": np.int64(curve_id), "n_inputs": np.int64(cs.num_inputs), "n_aux": np.int64(cs.num_aux),
           "r": H.ints_to_u64([r], 4), "s": H.ints_to_u64([s], 4),
           "z": H.fr_array(curve_id, cs.full_assignment()), "h": H.fr_array(curve_id, h, mont=False)}
    for which in "abc":
        ptr, cols, vals = cs.csr(which)
        out[which + "_ptr"] = np.asarray(ptr, dtype=np.uint32)
        out[which + "_col"] = np.asarray(cols, dtype=np.uint32)
        out[which + "_val"] = H.fr_array(curve_id, vals)
    for key, group in (("a_query", 1), ("b_g1_query", 1), ("b_g2_query", 2), ("h_query", 1), ("l_query", 1)):
        out[key + "_xy"], out[key + "_inf"] = H.points_array(curve_id, group, getattr(pk, key))
    out["g1_singles"] = H.points_array(curve_id, 1, [pk.alpha_g1, pk.beta_g1, pk.delta_g1])[0]
    out["g2_singles"] = H.points_array(curve_id, 2, [pk.beta_g2, pk.delta_g2])[0]
    for key, group, P in (("proof_a", 1, proof[0]), ("proof_b", 2, proof[1]), ("proof_c", 1, proof[2])):
        out[key + "_xy"], out[key + "_inf"] = H.points_array(curve_id, group, [P])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, "constraints", cs.num_constraints)


def msm_case(name, curve_id, group, n, seed):
    c = CURVES[(curve_id, group)]
    pts = H.multiples(curve_id, group, n)                   # P_i = (i+1) * G  
    pts[n // 3] = None                                      # one identity base
    sc = [stream_field(seed, i, c.r) for i in range(n)]
    sc[0], sc[1], sc[2] = 0, 1, c.r - 1
    res = c.to_affine(msm_pippenger(c, pts, sc, FR[curve_id].bits))
    xy, inf = H.points_array(curve_id, group, pts)
    rxy, rinf = H.points_array(curve_id, group, [res])
    np.savez_compressed(os.path.join(HERE, name + ".npz"), curve=np.int64(curve_id), group=np.int64(group), bases_xy=xy,
                        bases_inf=inf, scalars=H.ints_to_u64(sc, 4), result_xy=rxy, result_inf=rinf)
    print("wrote", name)


def ntt_case(name, curve_id, log_n, seed):
    fr = FR[curve_id]
    n = 1 << log_n
    vals = [stream_field(seed, i, fr.p) for i in range(n)]
    d = Domain(fr, n)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), curve=np.int64(curve_id), log_n=np.int64(log_n),
                        input=H.fr_array(curve_id, vals), fft=H.fr_array(curve_id, d.fft(vals)),
                        ifft=H.fr_array(curve_id, d.ifft(vals)), coset_fft=H.fr_array(curve_id, d.coset_fft(vals)),
                        coset_ifft=H.fr_array(curve_id, d.coset_ifft(vals)))
    print("wrote", name)


if __name__ == "__main__":
    only = sys.argv[1:]
    want = lambda n: not only or n in only
    for cid, tag in ((BLS12_381, "bls12_381"), (BN254, "bn254")):
        if want("ntt"):
            ntt_case("ntt_%s_2e8" % tag, cid, 8, 4)
        if want("msm"):
            msm_case("msm_%s_g1_256" % tag, cid, 1, 256, 3)
            msm_case("msm_%s_g2_64" % tag, cid, 2, 64, 3)
    if want("mini"):
        # groth16/tests/mini.rs:12-44 on the reference's curve
        groth16_case("groth16_mini_bls12_381", BLS12_381, lambda cs: mini_circuit(cs), 1)
        groth16_case("groth16_mini_bn254", BN254, lambda cs: mini_circuit(cs), 1)
    if want("mimc"):
        # BASELINE config 1: Groth16, BN256, 2^10-constraint MiMC chain
        groth16_case("groth16_mimc_bn254_2e10", BN254, lambda cs: mimc_circuit(cs, 1 << 10), 1)
        groth16_case("groth16_mimc_bls12_381_2e6", BLS12_381, lambda cs: mimc_circuit(cs, 1 << 6), 2)
