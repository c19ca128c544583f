"""CPU: host logic of the generic-skeleton variant (acinoset_b200/build.py, skeleton.py) - no GPU calls."""
import json

import numpy as np

from conftest import golden
from oracle import skel_fte


def _skel(tag="K1"):
    return json.loads(str(golden("generic_fk.npz")[tag + "_skeleton_json"]))


def test_path_masks_reproduce_the_builder_fk():
    """pose_r = root + sum of the link increments named by out_path[r] (incl. the twice-assigned hip1)."""
    from acinoset_b200 import skeleton

    for tag in ("K1", "K2"):
        skel = _skel(tag)
        flat = skeleton.flatten_skeleton(skel)
        f, names = skel_fte.pose_function(skel)
        assert names == flat["out_names"]
        x = np.array(golden("generic_fk.npz")[tag + "_x"][3], dtype=np.float64)
        L = len(flat["parts"])
        inc = []
        for l in range(len(flat["link_parent"])):
            a = flat["link_parent"][l]
            M = np.eye(3)
            m = flat["dof_mask"][a]
            if m & 2:
                M = skel_fte._rot(1, x[3 + L + a]) @ M
            if m & 1:
                M = skel_fte._rot(0, x[3 + a]) @ M
            if m & 4:
                M = skel_fte._rot(2, x[3 + 2 * L + a]) @ M
            inc.append((M.T if flat["link_flags"][l] else M) @ flat["link_tv"][l])
        pos = np.array([x[:3] + sum((inc[l] for l in range(len(inc)) if (int(flat["out_path"][r]) >> l) & 1), np.zeros(3))
                        for r in range(len(names))])
        assert np.abs(pos - f(x)).max() < 1e-14
        # hip2 -> hip1 overwrites shoulder1 -> hip1 (build.py:80): knee1 hangs off the second assignment
        k1, h2 = names.index("knee1"), names.index("hip2")
        assert int(flat["out_path"][k1]) & int(flat["out_path"][h2]) == int(flat["out_path"][h2])


def test_model_from_arrays_pairing_and_weights():
    from acinoset_b200 import build

    skel = _skel("K1")
    markers = list(skel["markers"])
    rng = np.random.default_rng(0)
    N, C = 4, 3
    meas = rng.normal(size=(N, C, len(markers), 2))
    lik = rng.uniform(0, 1, (N, C, len(markers)))
    K = np.tile(np.eye(3), (C, 1, 1))
    cams = (K, np.zeros((C, 4)), K.copy(), np.zeros((C, 3)))
    m = build.model_from_arrays(skel, cams, meas, lik, pair_by="index")
    names = m.flat["out_names"]
    assert m.P == 3 + 3 * 15 and m.N == N and m.x0.shape == (N, 48)
    for r in range(len(names)):
        if markers[r] == "neck":
            assert np.all(m.w[:, :, r] == 0)
        else:                                       # FK row r is compared with DLC marker markers[r] (by index)
            assert np.array_equal(m.meas[:, :, r], meas[:, :, r])
            assert np.array_equal(m.w[:, :, r] > 0, lik[:, :, r] > 0.4)
            assert set(np.unique(m.w[:, :, r])) <= {0.0, 1.0 / 3.0}
    mn = build.model_from_arrays(skel, cams, meas, lik, pair_by="name")
    for r, nm in enumerate(names):
        if nm == "neck":
            assert np.all(mn.w[:, :, r] == 0)
        else:
            assert np.array_equal(mn.meas[:, :, r], meas[:, :, markers.index(nm)])
    lo, hi = build.bounds(15)
    lo2, hi2 = skel_fte.bounds(15)
    assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2)
    assert lo[2] == -np.pi / 2 and np.isinf(lo[0]) and np.isinf(lo[3 * 15 - 1]) and lo[3 * 15 - 2] == -np.pi / 2


def test_redescending_loss_name_matches_reference_vectors():
    from acinoset_b200 import build

    g = golden("loss.npz")
    assert np.abs(build.redescending_loss(g["e"], 3, 10, 20) - g["rho_3_10_20"]).max() < 1e-12


def test_reference_module_names_are_importable():
    """A user of the reference finds the names its scripts call (no GPU needed to import)."""
    from acinoset_b200 import all_optimizations, app, build, calib, misc, sba, stereo, utils

    for mod, names in [
        (calib, ["project_points_fisheye", "triangulate_points_fisheye", "get_pairwise_3d_points_from_df", "project_points",
                 "triangulate_points", "create_undistort_point_function", "bundle_adjust_points_and_extrinsics",
                 "bundle_adjust_points_only", "bundle_adjust_board_points_and_extrinsics",
                 "create_bundle_adjustment_jacobian_sparsity_matrix", "prepare_calib_board_data_for_bundle_adjustment",
                 "cost_func_points_extrinsics", "params_to_points_extrinsics"]),
        (stereo, ["calibrate_pair_extrinsics_fisheye", "calibrate_pair_extrinsics", "calibrate_pairwise_extrinsics"]),
        (app, ["calibrate_fisheye_extrinsics_pairwise", "calibrate_standard_extrinsics_pairwise", "sba_board_points_fisheye",
               "sba_points_fisheye", "save_tri", "save_sba", "save_ekf", "save_fte", "save_optimised_cheetah", "save_3d_cheetah_as_2d"]),
        (misc, ["get_markers", "get_pose_params", "get_3d_marker_coords", "redescending_loss", "rot_x", "rot_y", "rot_z"]),
        (all_optimizations, ["fte", "ekf", "sba", "tri", "pt3d_to_2d", "pt3d_to_x2d", "pt3d_to_y2d"]),
        (build, ["load_skeleton", "build_model", "solve_optimisation", "save_data", "convert_to_dict", "redescending_loss",
                 "pt3d_to_2d", "np_rot_x"]),
        (utils, ["load_scene", "save_scene", "load_camera", "save_camera", "load_points", "save_points", "find_scene_file",
                 "load_dlc_points_as_df", "create_dlc_points_2d_file", "create_board_object_pts"]),
        (sba, ["sba_board_points_fisheye", "sba_points_fisheye", "jac_points_extrinsics"]),
    ]:
        for n in names:
            assert callable(getattr(mod, n)), (mod.__name__, n)
    assert len(misc.get_markers()) == 20 and misc.get_markers()[2] == "nose"
    assert misc.get_pose_params()["psi_0"] == 5
    import numpy as np

    assert np.allclose(misc.rot_x(0.3) @ misc.rot_x(0.3).T, np.eye(3))


def test_sparsity_matrix_matches_reference_golden():
    """create_bundle_adjustment_jacobian_sparsity_matrix against the pattern the reference's own function produced
    (tests/golden/sba.npz A_rows / A_cols, calib.py:196-207), for the extrinsics and the points-only layout."""
    from conftest import golden
    from oracle import sba as osba
    from acinoset_b200 import sba

    g = golden("sba.npz")
    for tag in ("static", "rotating"):
        n_pts = len(g[f"{tag}_points_3d"])
        A = sba.create_bundle_adjustment_jacobian_sparsity_matrix(2, 6, g[f"{tag}_cidx"], n_pts, g[f"{tag}_pidx"])
        assert A.shape == tuple(g[f"{tag}_A_shape"]) and A.dtype == int
        r, c = A.tocoo().row, A.tocoo().col
        o = np.lexsort((c, r))
        o2 = np.lexsort((g[f"{tag}_A_cols"], g[f"{tag}_A_rows"]))
        assert np.array_equal(r[o], g[f"{tag}_A_rows"][o2]) and np.array_equal(c[o], g[f"{tag}_A_cols"][o2])
        assert A.tocsr().max() == 1
        A0 = sba.create_bundle_adjustment_jacobian_sparsity_matrix(2, 0, g[f"{tag}_cidx"], n_pts, g[f"{tag}_pidx"])
        assert np.array_equal(A0.toarray(), osba.sparsity(2, 0, g[f"{tag}_cidx"], n_pts, g[f"{tag}_pidx"]))
