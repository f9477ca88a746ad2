"""bindings/ (SURVEY.md 8f-1): the generated Rust declarations are current, the C ABI works from plain C through dlopen
(tests/c/abi_dlopen.c) with a byte-layout fixture (ark-serialize compressed points -> zkb_points_decompress ->
zkb_srs_upload -> zkb_msm), and zkb_points_decompress matches the oracle's serializer on all four groups."""
import os
import random
import re
import struct
import subprocess
import sys

import numpy as np
import pytest

from oracle.pyref import serialize as S
from oracle.pyref.curves import CURVES
from oracle.pyref.fields import BLS12_381, BN254, FQ
from oracle.pyref.msm import msm_naive
from tests import helpers as H

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "ckb_zkp_b200", "libzkb.so")


def test_ffi_rs_is_generated_from_the_header():
    rc = subprocess.call([sys.executable, os.path.join(ROOT, "tools", "gen_zkb_sys.py"), "--check"])
    assert rc == 0, "bindings/zkb-sys/src/ffi.rs is stale: run python tools/gen_zkb_sys.py"
    ffi = open(os.path.join(ROOT, "bindings", "zkb-sys", "src", "ffi.rs")).read()
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "zkb.h")).read(), flags=re.S)
    declared = set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", hdr))
    bound = set(re.findall(r"pub fn (zkb_\w+)\(", ffi))
    assert declared == bound, declared ^ bound


def test_patches_touch_the_cited_call_sites():
    """each patch names the reference file it applies to and the glue module it includes exists"""
    pdir = os.path.join(ROOT, "bindings", "patches")
    for name, target, glue in (("zkp-groth16.diff", "groth16/src/prover.rs", "groth16_src_zkb_backend.rs"),
                               ("zkp-marlin-kzg10.diff", "marlin/src/pc/kzg10.rs", "marlin_src_pc_zkb_backend.rs"),
                               ("zkp-curve.diff", "curve/src/lib.rs", "curve_src_zkb_backend.rs")):
        text = open(os.path.join(pdir, name)).read()
        assert ("+++ b/" + target) in text, name
        assert "zkb_backend" in text, name
        glue_src = open(os.path.join(pdir, "new_files", glue)).read()
        assert "zkb_sys" in glue_src, glue
        for fn in re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", glue_src):          # the glue only calls what the header declares
            assert re.search(r"\b%s\s*\(" % fn, open(os.path.join(ROOT, "include", "zkb.h")).read()), (glue, fn)


def _build_c(tmp_path):
    exe = str(tmp_path / "abi_dlopen")
    subprocess.check_call(["gcc", "-O1", "-Wall", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "c", "abi_dlopen.c"), "-ldl"])
    hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "zkb.h")).read(), flags=re.S)
    syms = sorted(set(re.findall(r"\b(zkb_[a-z0-9_]+)\s*\(", hdr)))
    (tmp_path / "symbols.txt").write_text("\n".join(syms) + "\n")
    return exe, str(tmp_path / "symbols.txt"), len(syms)


def _fixture(path, cid, group, n=24, seed=5):
    """compressed points (one identity among them) + the limbs / MSM result the C program must find"""
    c = CURVES[(cid, group)]
    rng = random.Random(seed)
    pts = [c.mul_affine(c.gen, rng.randrange(1, c.r)) for _ in range(n - 1)]
    pts.insert(3, None)
    comp = b"".join(S.compress(cid, group, P) for P in pts)
    xy, inf = H.points_array(cid, group, pts)
    xy[inf == 1] = 0                                   # the device's identity placeholder
    scalars = [rng.randrange(c.r) for _ in range(n)]
    scalars[0], scalars[1] = 0, c.r - 1
    want = c.to_affine(msm_naive(c, pts, scalars))
    wxy, winf = H.points_array(cid, group, [want])
    words = xy.shape[1]
    with open(path, "wb") as f:
        f.write(struct.pack("<4I", cid, group, n, words))
        f.write(comp)
        f.write(xy.tobytes())
        f.write(inf.tobytes())
        f.write(H.ints_to_u64(scalars, 4).tobytes())
        f.write(wxy.tobytes())
        f.write(winf.tobytes())


def test_c_program_resolves_every_symbol(tmp_path):
    """runs everywhere: without a GPU the program must stop at zkb_init with ZKB_E_NO_DEVICE (no CPU fallback)"""
    exe, syms, n_syms = _build_c(tmp_path)
    fx = str(tmp_path / "fx.bin")
    _fixture(fx, BN254, 1, n=6)
    out = subprocess.run([exe, LIB, syms, fx], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert "symbols %d" % n_syms in out.stdout and "keccak ok" in out.stdout
    import torch
    assert ("gpu ok" if torch.cuda.is_available() else "no-device") in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("cid,group", [(BN254, 1), (BLS12_381, 1), (BLS12_381, 2)])
def test_c_program_byte_layout_fixture(tmp_path, cid, group):
    exe, syms, _ = _build_c(tmp_path)
    fx = str(tmp_path / "fx.bin")
    _fixture(fx, cid, group)
    out = subprocess.run([exe, LIB, syms, fx], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr
    assert "decompress ok" in out.stdout and "msm ok" in out.stdout and "gpu ok" in out.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("cid,group", [(BN254, 1), (BN254, 2), (BLS12_381, 1), (BLS12_381, 2)])
def test_points_decompress_matches_oracle(ctx, cid, group):
    c = CURVES[(cid, group)]
    rng = random.Random(100 + 10 * cid + group)
    pts = H.multiples(cid, group, 300, start=rng.randrange(1, 1 << 60))
    pts[7] = None
    pts[11] = c.neg_affine(pts[11])
    data = bytearray(b"".join(S.compress(cid, group, P) for P in pts))
    stride = len(data) // len(pts)
    # rejections: a non-canonical x, an x with no point, and (where the cofactor is not 1) a point outside the subgroup
    p = FQ[cid].p
    nb = 8 * FQ[cid].limbs
    bad = {}
    big = (p + 1).to_bytes(nb, "little")
    if big[-1] & 0xC0 == 0:
        data[20 * stride:21 * stride] = big * (2 if group == 2 else 1)
        bad[20] = 1
    x = 1
    while True:
        enc = (int(x).to_bytes(nb, "little") + bytes(nb)) if group == 2 else int(x).to_bytes(nb, "little")
        try:
            S.decompress(cid, group, enc)
            x += 1
        except ValueError:
            break
    data[30 * stride:31 * stride] = enc
    bad[30] = 2
    xy, inf, status = ctx.points_decompress(cid, group, np.frombuffer(bytes(data), dtype=np.uint8), check_subgroup=True)
    for i, P in enumerate(pts):
        if i in bad:
            assert status[i] == bad[i] and inf[i] == 1, i
            continue
        assert status[i] == 0, i
        assert H.array_point(cid, group, xy[i], inf[i]) == P, i
    # without the subgroup check a curve point of the wrong order decodes fine, with it the point is rejected
    if (cid, group) != (BN254, 1):
        x = 2
        while True:
            enc = (int(x).to_bytes(nb, "little") + bytes(nb)) if group == 2 else int(x).to_bytes(nb, "little")
            try:
                P = S.decompress(cid, group, enc)
                if P is not None and c.to_affine(c.mul(c.from_affine(P), c.r)) is not None:
                    break
            except ValueError:
                pass
            x += 1
        one = np.frombuffer(enc, dtype=np.uint8)
        xy1, inf1, st1 = ctx.points_decompress(cid, group, one, check_subgroup=False)
        assert st1[0] == 0 and H.array_point(cid, group, xy1[0], inf1[0]) == P
        _, inf2, st2 = ctx.points_decompress(cid, group, one, check_subgroup=True)
        assert st2[0] == 3 and inf2[0] == 1


@pytest.mark.gpu
def test_round2_entry_points_accept_empty_and_reject_bad_arguments(ctx):
    """empty inputs are no-ops, malformed ones are ZKB_E_INVALID with a message (never a crash)"""
    from ckb_zkp_b200 import _lib
    from ckb_zkp_b200.backend import ZkbError
    xy, inf, st = ctx.points_decompress(BLS12_381, 1, np.zeros((0, 48), dtype=np.uint8))
    assert xy.shape == (0, 12) and inf.shape == (0,) and st.shape == (0,)
    assert ctx.fr_prefix_product(BN254, np.zeros((0, 4), dtype=np.uint64)).shape == (0, 4)
    one = H.fr_array(BN254, [7])
    assert H.fr_ints(BN254, ctx.fr_prefix_product(BN254, one)) == [1]            # out[0] = 1 whatever the input
    assert ctx.msm_batch([], []) == []
    g1 = H.points_array(BN254, 1, H.multiples(BN254, 1, 4))
    g2 = H.points_array(BN254, 2, H.multiples(BN254, 2, 4))
    s1, s2 = ctx.srs_upload(BN254, 1, *g1), ctx.srs_upload(BN254, 2, *g2)
    sc = H.fr_array(BN254, [1, 2, 3, 4])
    with pytest.raises(ZkbError):                                             # mixed groups in one batch
        ctx.msm_batch([s1, s2], [sc, sc])
    got = ctx.msm_batch([s1], [sc[:0]])                                          # zero scalars: the identity
    assert got[0][1] is True
    rc = ctx.lib.zkb_points_decompress(ctx.handle, 7, 1, None, 1, 0, None, None, None)
    assert rc == _lib.E_INVALID
    # zkb_multi_pairing / zkb_poly_eval_batch: no groups is a no-op, empty groups / unknown curves / null buffers are errors
    import ctypes
    assert ctx.multi_pairing(BN254, (np.zeros((0, 8), dtype=np.uint64), None), (np.zeros((0, 16), dtype=np.uint64), None), 3).shape == (0, 48)
    gt = np.zeros(48, dtype=np.uint64)
    p1, p2 = np.ascontiguousarray(g1[0][:1]), np.ascontiguousarray(g2[0][:1])
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert ctx.lib.zkb_multi_pairing(ctx.handle, BN254, vp(p1), None, vp(p2), None, 1, 0, vp(gt)) == _lib.E_INVALID
    assert ctx.lib.zkb_multi_pairing(ctx.handle, 9, vp(p1), None, vp(p2), None, 1, 1, vp(gt)) == _lib.E_INVALID
    assert ctx.lib.zkb_multi_pairing(ctx.handle, BN254, None, None, vp(p2), None, 1, 1, vp(gt)) == _lib.E_INVALID
    assert ctx.lib.zkb_multi_pairing(ctx.handle, BN254, vp(p1), None, vp(p2), None, 1, 1, vp(gt)) == 0 and gt.any()
    assert ctx.lib.zkb_poly_eval_batch(ctx.handle, BN254, 1, None, None, None, None) == _lib.E_INVALID
    assert ctx.lib.zkb_poly_eval_batch(ctx.handle, BN254, 0, None, None, None, None) == 0
    s1.free()
    s2.free()
