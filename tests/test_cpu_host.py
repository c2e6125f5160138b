"""CPU tests of the host-side logic and of the C-ABI surface (no GPU work)."""
import ctypes as C
import re
import struct
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_c_abi_library_exports_every_declared_symbol():
    """libacmmp_b200.so loads and exports exactly what include/acmmp_b200.h declares."""
    import acmmp_b200
    header = (ROOT / "include" / "acmmp_b200.h").read_text()
    declared = set(re.findall(r"\b(acmmp_[a-z0-9_]+)\s*\(", header))
    declared -= {"acmmp_ctx", "acmmp_camera", "acmmp_params"}
    lib = acmmp_b200.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert set(acmmp_b200.exported_symbols()) == declared
    assert lib.acmmp_abi_sizeof_camera() == 120 and lib.acmmp_abi_sizeof_params() == 68
    assert b"sm_100a" in lib.acmmp_version()


def test_default_params_are_the_reference_defaults():
    import acmmp_b200
    p = acmmp_b200.Params()
    acmmp_b200.lib().acmmp_default_params(C.byref(p))
    assert (p.max_iterations, p.patch_size, p.radius_increment, p.top_k, p.max_image_size) == (3, 11, 2, 4, 3200)
    assert (p.sigma_spatial, p.sigma_color) == (5.0, 3.0)
    assert not (p.geom_consistency or p.planar_prior or p.hierarchy or p.upsample)


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device the product refuses to run; it never routes through the oracle."""
    import torch
    import acmmp_b200
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(acmmp_b200.AcmmpError):
        acmmp_b200.Context(0)
    src = "".join(p.read_text() for p in (ROOT / "acmmp-spherical_b200").rglob("*.py"))
    src += "".join(p.read_text() for p in (ROOT / "acmmp-spherical_b200" / "csrc").glob("*"))
    # nothing in the product tree imports, loads or names the checkers (oracle/, libacmmp_ref, libacmmp_oracle)
    import re
    hits = [m.group(0) for m in re.finditer(r"(?im)^.*\b(?:import|from|include|CDLL|dlopen)\b.*oracle.*$", src)]
    assert not hits, hits
    assert "cpu_oracle" not in src and "acmmp_oracle" not in src and "ref_driver" not in src and "libacmmp_ref" not in src


def test_pyramid_schedule_matches_reference_rules():
    from acmmp_b200.pipeline import pyramid_sizes
    assert pyramid_sizes(3200, 2130) == [800, 1600, 3200]          # main.cpp:35-71, :420-425
    assert pyramid_sizes(4096, 2048) == [800, 1600, 3200]          # capped at 3200
    assert pyramid_sizes(6000, 4000) == [800, 1600, 3200]
    assert pyramid_sizes(640, 480) == [640]
    assert pyramid_sizes(1001, 700) == [500, 1001]


def test_scale_problem_rounds_like_input_initialization():
    from acmmp_b200 import synth
    sc = synth.make_pinhole_scene(n_views=2, width=200, height=133, focal=150.0, seed=3)
    imgs, cams = synth.scale_problem(sc.images, sc.cams, 100)
    assert imgs[0].shape == (66, 100) or imgs[0].shape == (67, 100)       # round(133 * 0.5)
    assert abs(cams[0].K[0] - 150.0 * (100 / 200)) < 1e-4
    assert abs(cams[0].K[5] - sc.cams[0].K[5] * (imgs[0].shape[0] / 133)) < 1e-3
    same, _ = synth.scale_problem(sc.images, sc.cams, 4000)
    assert same[0] is sc.images[0]


def test_dmb_roundtrip(tmp_path):
    from oracle.ref_driver import read_dmb, write_dmb
    a = np.random.default_rng(0).random((7, 5), dtype=np.float32)
    n = np.random.default_rng(1).random((7, 5, 3), dtype=np.float32)
    write_dmb(tmp_path / "d.dmb", a)
    write_dmb(tmp_path / "n.dmb", n)
    raw = (tmp_path / "d.dmb").read_bytes()
    assert struct.unpack("<iiii", raw[:16]) == (1, 7, 5, 1)              # ACMMP.cpp:395-420
    assert np.array_equal(read_dmb(tmp_path / "d.dmb"), a)
    assert np.array_equal(read_dmb(tmp_path / "n.dmb"), n)


def test_planar_prior_stage_on_a_plane():
    """Support points -> Delaunay -> per-triangle plane: a fronto-parallel plane must come back."""
    from acmmp_b200 import synth
    from acmmp_b200.prior import planar_prior, support_points
    sc = synth.make_pinhole_scene(n_views=2, width=160, height=120, focal=120.0, seed=5)
    H, W = 120, 160
    depth = np.full((H, W), 3.0, np.float32)
    costs = np.full((H, W), 0.05, np.float32)
    costs[::2, ::2] = 0.01
    pts = support_points(costs)
    assert len(pts) == (W // 5) * (H // 5)
    params, masks = planar_prior(sc.cams[0], depth, costs, 1.0, 10.0)
    assert params.shape[1] == 4 and masks.max() == len(params)
    covered = masks > 0
    assert covered.mean() > 0.8
    # z = 3 in the camera frame: n = (0, 0, -1) up to sign convention of the reference (w >= 0), d = 3
    n = params[:, :3]
    assert np.allclose(np.abs(n[:, 2]), 1.0, atol=1e-3) and np.allclose(np.abs(params[:, 3]), 3.0, atol=1e-2)


def test_on_disk_contract_is_what_the_reference_reads(tmp_path):
    from acmmp_b200 import synth
    sc = synth.make_pinhole_scene(n_views=3, width=64, height=48, focal=60.0, seed=7)
    synth.write_dense_folder(sc, str(tmp_path))
    txt = (tmp_path / "cams" / "00000001_cam.txt").read_text().split()
    assert txt[0] == "extrinsic" and txt[17] == "intrinsic" and len(txt) == 18 + 9 + 4
    pair = (tmp_path / "pair.txt").read_text().split()
    assert pair[0] == "3" and pair[1] == "0" and pair[2] == "2"
    ss = synth.make_sphere_scene(n_views=2, width=64, height=32, seed=8)
    synth.write_dense_folder(ss, str(tmp_path / "s"))
    txt = (tmp_path / "s" / "cams" / "00000000_cam.txt").read_text().split()
    assert txt[18] == "SPHERE" and len(txt) == 18 + 1 + 3 + 4
