"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE'S OWN CODE.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py [--skip-sympy-jac]

What is executed:
  * /root/reference/src/calib/{calib,utils}.py imported UNMODIFIED through oracle.ref_shim
    (stub nptyping, np.float/np.int aliases, stub matplotlib)  -> projection, triangulation,
    TRI driver, SBA assembly + residuals (OpenCV 4.13.0 / SciPy 1.18.1 of this image).
  * source text of src/all_optimizations.py:66-190 (cheetah FK, SymPy) and :193-209
    (pt3d_to_2d), src/build.py:32-95 (generic skeleton builder) and :382-395 (redescending
    loss) exec'd verbatim with NumPy/SymPy intrinsics in place of Pyomo's (pyomo is not
    installable here); nothing is re-typed.
  * shipped artefacts: data/results/traj_results.pickle, data/old_results/run1.pickle,
    skeletons/*.pickle, data/sunday_amelia/extrinsic_calib/* (K1-K6 of SURVEY.md section 4).
Outputs are small .npz / .json files committed next to this script.
"""
import argparse
import json
import os
import pickle
import sys
import textwrap

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
import synth  # noqa: E402

REF = ref_shim.REF_ROOT


def ref_src(rel, a, b):
    return textwrap.dedent(ref_shim.reference_source_lines(rel, a, b))


def gen_cheetah_fk(skip_jac):
    import sympy as sp

    ns = dict(sp=sp, np=np, sin=np.sin, cos=np.cos)
    exec(ref_src("src/all_optimizations.py", 66, 190), ns)
    pose_to_3d, positions, sym_list = ns["pose_to_3d"], ns["positions"], ns["sym_list"]
    fs = positions.free_symbols
    active = [i for i, s in enumerate(sym_list) if s in fs]
    rng = np.random.default_rng(20260101)
    x45 = np.zeros((48, 45))
    x45[:, active] = rng.normal(0, 1.0, (48, len(active)))
    x45[:, :3] = rng.uniform([-2, 3, 0.2], [6, 10, 1.2], (48, 3))
    # a non-active slot must not matter: give some rows junk there
    inactive = [i for i in range(45) if i not in active]
    x45[40:, inactive] = rng.normal(0, 1, (8, len(inactive)))
    pos = np.array([pose_to_3d(*x) for x in x45])
    out = dict(x45=x45, positions=pos, active=np.array(active))
    if not skip_jac:
        Jsym = positions.reshape(60, 1).jacobian([sym_list[i] for i in active])
        Jf = sp.lambdify(sym_list, Jsym, modules=[{"sin": np.sin, "cos": np.cos, "ImmutableDenseMatrix": np.array}])
        out["jac"] = np.array([Jf(*x) for x in x45[:12]]).reshape(12, 20, 3, len(active))
    np.savez_compressed(os.path.join(HERE, "cheetah_fk.npz"), **out)
    print("cheetah_fk", pos.shape, "active", active)


def gen_fisheye(calib, utils):
    import cv2

    K, D, R, t, res = utils.load_scene(os.path.join(REF, "configs", "dummy_scene.json"))
    ns = dict(np=np, atan=np.arctan)
    exec(ref_src("src/all_optimizations.py", 193, 209), ns)
    pt3d_to_2d = ns["pt3d_to_2d"]
    rng = np.random.default_rng(7)
    X = rng.uniform([-3, 2, 0], [7, 11, 1.5], (96, 3))
    uv_sym = np.zeros((6, 96, 2))
    uv_cv = np.zeros((6, 96, 2))
    for c in range(6):
        Dc = D[c].reshape(4)
        for i in range(96):
            uv_sym[c, i] = pt3d_to_2d(X[i, 0], X[i, 1], X[i, 2], K[c], Dc, R[c], t[c])
        uv_cv[c] = calib.project_points_fisheye(X, K[c], D[c], R[c], t[c])
    # undistort: nominal + wide angle + a nastier distortion (thursday_kiara cam 2)
    pts_in = rng.uniform([-1500, -1200], [4200, 2700], (400, 2))
    und_dummy = cv2.fisheye.undistortPoints(pts_in.reshape(-1, 1, 2), K[1], D[1]).reshape(-1, 2)
    kk = json.load(open(os.path.join(REF, "data", "thursday_kiara", "extrinsic_calib", "4_cam_scene.json")))
    K2 = np.array(kk["cameras"][1]["k"])
    D2 = np.array(kk["cameras"][1]["d"]).reshape(4, 1)
    und_kiara = cv2.fisheye.undistortPoints(pts_in.reshape(-1, 1, 2), K2, D2).reshape(-1, 2)
    # Rodrigues both ways + jacobian
    rv = rng.normal(0, 1, (8, 3))
    rv[0] = 0
    rv[1] *= 1e-9
    Rm = np.array([cv2.Rodrigues(r)[0] for r in rv])
    Rj = np.array([cv2.Rodrigues(r)[1] for r in rv])  # (8,3,9): d R.flat / d rvec[k] at [k, :]
    rv_back = np.array([cv2.Rodrigues(m)[0].ravel() for m in Rm])
    np.savez_compressed(os.path.join(HERE, "fisheye.npz"), K=K, D=D, R=R, t=t, res=np.array(res), X=X,
                        uv_pt3d_to_2d=uv_sym, uv_cv2=uv_cv, und_in=pts_in, und_dummy=und_dummy,
                        K_kiara=K2, D_kiara=D2, und_kiara=und_kiara, rvec=rv, rmat=Rm, rjac=Rj,
                        rvec_back=rv_back)
    print("fisheye: max |pt3d_to_2d - cv2| =", np.abs(uv_sym - uv_cv).max(),
          "sentinels:", int((und_dummy[:, 0] == -1e6).sum()), int((und_kiara[:, 0] == -1e6).sum()))


def gen_loss():
    ns = dict(np=np)
    exec(ref_src("src/build.py", 382, 395), ns)
    f = ns["redescending_loss"]
    e = np.concatenate([np.linspace(-60, 60, 961), [0.0, 1e-9, 0.2, 400.0, -400.0]])
    rho = np.array([f(v, 3, 10, 20) for v in e])
    rho2 = np.array([f(v, 3, 5, 15) for v in e])  # the commented variant at build.py:298
    np.savez_compressed(os.path.join(HERE, "loss.npz"), e=e, rho_3_10_20=rho, rho_3_5_15=rho2)
    print("loss", rho[:3])


def gen_triangulate(calib, utils):
    from oracle import fisheye as ofe, skeleton as osk

    K, D, R, t, res = utils.load_scene(os.path.join(REF, "configs", "dummy_scene.json"))
    cams = (K, D.reshape(-1, 4), R, t.reshape(-1, 3), res)
    # config 1 of BASELINE.json: 6 cam x 20 kpt x 90 frames
    prob = synth.make_fte_problem(90, osk.cheetah_fk_active, ofe.project, seed=1, cams=cams)
    df = synth.dense_to_long_df(prob["meas"], prob["lik"], osk.MARKERS)
    dff = df[df["likelihood"] > 0.5].reset_index(drop=True)
    out = calib.get_pairwise_3d_points_from_df(dff, K, D.reshape(-1, 4), R, t, calib.triangulate_points_fisheye)
    pos = np.full((90, 20, 3), np.nan)
    mi = {m: i for i, m in enumerate(osk.MARKERS)}
    for fr, mk, x, y, z in out[["frame", "marker", "x", "y", "z"]].values:
        pos[int(fr), mi[mk]] = (x, y, z)
    # noise-free variant: DLT exactness (K8)
    P = osk.cheetah_fk_active(prob["x_true"])
    rng = np.random.default_rng(3)
    meas0, lik0 = synth.make_measurements(P, cams, ofe.project, rng, noise_px=0.0, outlier_frac=0.0, low_lik_frac=0.0)
    df0 = synth.dense_to_long_df(meas0, lik0, osk.MARKERS)
    out0 = calib.get_pairwise_3d_points_from_df(df0[df0["likelihood"] > 0.5].reset_index(drop=True), K,
                                                D.reshape(-1, 4), R, t, calib.triangulate_points_fisheye)
    pos0 = np.full((90, 20, 3), np.nan)
    for fr, mk, x, y, z in out0[["frame", "marker", "x", "y", "z"]].values:
        pos0[int(fr), mi[mk]] = (x, y, z)
    # raw pair call with noisy + wide points
    a = prob["meas"][:, 2].reshape(-1, 2)[:300]
    b = prob["meas"][:, 3].reshape(-1, 2)[:300]
    pair = calib.triangulate_points_fisheye(a, b, K[2], D[2], R[2], t[2], K[3], D[3], R[3], t[3])
    np.savez_compressed(os.path.join(HERE, "triangulate.npz"), meas=prob["meas"], lik=prob["lik"],
                        x_true=prob["x_true"], tri_pos=pos, meas_clean=meas0, lik_clean=lik0, tri_pos_clean=pos0,
                        P_true=P, pair_a=a, pair_b=b, pair_out=pair)
    ok = ~np.isnan(pos0[..., 0])
    print("triangulate: rows", len(out), "clean max err", np.abs(pos0[ok] - P[ok]).max())


def gen_generic_fk():
    import sympy as sp

    out = {}
    for tag, skel_f, res_f in [("K1", "skeletons/new_human.pickle", "data/results/traj_results.pickle"),
                               ("K2", "skeletons/human.pickle", "data/old_results/run1.pickle")]:
        skel = pickle.load(open(os.path.join(REF, skel_f), "rb"))
        res = pickle.load(open(os.path.join(REF, res_f), "rb"))
        ns = dict(sp=sp, np=np, sin=np.sin, cos=np.cos, skel_dict=skel, print=lambda *a, **k: None)
        exec(ref_src("src/build.py", 399, 424), ns)   # rot_x / rot_y / rot_z (SymPy)
        exec(ref_src("src/build.py", 32, 95), ns)     # the builder body
        f = ns["pose_to_3d"]
        x = np.asarray(res["x"], dtype=np.float64)
        pos_exec = np.array([f(*row) for row in x])
        out[tag + "_skeleton_json"] = np.array(json.dumps(skel))
        out[tag + "_x"] = x
        out[tag + "_dx"] = np.asarray(res["dx"])
        out[tag + "_ddx"] = np.asarray(res["ddx"])
        out[tag + "_positions_pickle"] = np.asarray(res["positions"])
        out[tag + "_positions_exec"] = pos_exec
        out[tag + "_pose_order"] = np.array(list(ns["pose_dict"].keys()))
        print("generic_fk", tag, "exec vs shipped:", np.abs(pos_exec - np.asarray(res["positions"])).max())
    np.savez_compressed(os.path.join(HERE, "generic_fk.npz"), **out)


def gen_sba(calib, utils):
    from scipy.sparse import issparse

    base = os.path.join(REF, "data", "sunday_amelia", "extrinsic_calib")
    out = {}
    for tag, cams in [("static", (3, 4)), ("rotating", (1, 2))]:
        K, D, R, t, res = utils.load_scene(os.path.join(base, f"4_cam_scene_{tag}.json"))
        Ks, Ds, Rs, ts, _ = utils.load_scene(os.path.join(base, f"4_cam_scene_{tag}_sba.json"))
        pts, fns = [], []
        for c in cams:
            d = json.load(open(os.path.join(base, "points", f"points_cam{c}.json")))
            fns.append(list(d["points"].keys()))
            pts.append(np.array(list(d["points"].values()), dtype=np.float32))
            board_shape = tuple(d["board_shape"])
        # deterministic view order: the reference iterates a set; force sorted order by
        # handing it pre-sorted, identical name lists (cost is order-invariant)
        p2d, p3d, pidx, cidx = calib.prepare_calib_board_data_for_bundle_adjustment(
            pts, fns, board_shape, K, D, R, t, calib.triangulate_points_fisheye)
        n_cam, n_pts = len(K), len(p3d)
        r_vecs = np.array([__import__("cv2").Rodrigues(r)[0] for r in R], dtype=np.float64).flatten()
        x0 = np.concatenate([r_vecs, t.flatten(), p3d.flatten()])
        f0 = calib.cost_func_points_extrinsics(x0, n_cam, n_pts, pidx, cidx, K, D, p2d, calib.project_points_fisheye)
        A = calib.create_bundle_adjustment_jacobian_sparsity_matrix(n_cam, 6, cidx, n_pts, pidx)
        A = A.toarray() if issparse(A) else np.asarray(A)
        rows, cols = np.nonzero(A)
        out.update({
            f"{tag}_K": K, f"{tag}_D": D, f"{tag}_R": R, f"{tag}_t": t,
            f"{tag}_R_sba": Rs, f"{tag}_t_sba": ts,
            f"{tag}_img_pts_a": pts[0], f"{tag}_img_pts_b": pts[1],
            f"{tag}_fnames_a": np.array(fns[0]), f"{tag}_fnames_b": np.array(fns[1]),
            f"{tag}_points_2d": p2d, f"{tag}_points_3d": p3d, f"{tag}_pidx": pidx, f"{tag}_cidx": cidx,
            f"{tag}_x0": x0, f"{tag}_f0": f0, f"{tag}_A_rows": rows.astype(np.int32),
            f"{tag}_A_cols": cols.astype(np.int32), f"{tag}_A_shape": np.array(A.shape),
        })
        cost0 = 0.5 * np.sum(np.log1p(f0 ** 2))
        print("sba", tag, "n_obs", len(pidx), "n_pts", n_pts, "initial cauchy cost", cost0)
        out[f"{tag}_cost0"] = cost0
    out["board_shape"] = np.array(board_shape)
    np.savez_compressed(os.path.join(HERE, "sba.npz"), **out)


def gen_pinhole(calib, utils):
    """project_points / create_undistort_point_function / triangulate_points (calib.py:25-30,52-66) run through
    the reference's own functions with the rational (8-coefficient, calib.py:18), plumb-bob (5) and
    thin-prism (12) coefficient sets."""
    import cv2

    rng = np.random.default_rng(11)
    K1 = np.array([[1210.0, 0, 1352.0], [0, 1198.0, 761.0], [0, 0, 1]])
    K2 = np.array([[1180.0, 0, 1330.0], [0, 1185.0, 770.0], [0, 0, 1]])
    dists = {
        "d5": np.array([-0.21, 0.07, 0.0012, -0.0008, -0.011]),
        "d8": np.array([0.11, -0.05, 0.001, -0.002, 0.01, 0.02, -0.01, 0.003]),
        "d12": np.array([0.08, -0.03, 0.0007, 0.0011, 0.004, 0.015, -0.006, 0.001, 0.0009, -0.0004, 0.0006, 0.0003]),
    }
    r1 = cv2.Rodrigues(np.array([0.1, -0.2, 0.05]))[0]
    t1 = np.array([[0.1], [0.2], [0.3]])
    r2 = cv2.Rodrigues(np.array([-0.05, 0.35, 0.02]))[0]
    t2 = np.array([[-1.4], [0.1], [0.5]])
    X = rng.normal(0, 0.8, (200, 3)) + [0, 0, 5.0]
    out = dict(K1=K1, K2=K2, r1=r1, t1=t1, r2=r2, t2=t2, X=X)
    for tag, d in dists.items():
        uv1 = calib.project_points(X, K1, d, r1, t1)
        uv2 = calib.project_points(X, K2, d, r2, t2)
        uv1_rvec = calib.project_points(X, K1, d, cv2.Rodrigues(r1)[0], t1)
        und = calib.create_undistort_point_function(K1, d)(uv1.astype(np.float64))
        noisy1 = uv1 + rng.normal(0, 1.0, uv1.shape)
        noisy2 = uv2 + rng.normal(0, 1.0, uv2.shape)
        tri = calib.triangulate_points(noisy1, noisy2, K1, d, r1, t1, K2, d, r2, t2)
        tri0 = calib.triangulate_points(uv1, uv2, K1, d, r1, t1, K2, d, r2, t2)
        out.update({f"{tag}": d, f"{tag}_uv1": uv1, f"{tag}_uv2": uv2, f"{tag}_uv1_rvec": uv1_rvec, f"{tag}_und": und,
                    f"{tag}_noisy1": noisy1, f"{tag}_noisy2": noisy2, f"{tag}_tri": tri, f"{tag}_tri_clean": tri0})
        print("pinhole", tag, "clean triangulation max err", np.abs(tri0 - X).max())
    np.savez_compressed(os.path.join(HERE, "pinhole.npz"), **out)


def gen_stereo(calib, utils):
    """Pairwise extrinsic calibration (calib.py:125-134): cv2.fisheye.stereoCalibrate called exactly as the reference
    calls it (CALIB_FIX_INTRINSIC, 100 iterations, eps 1e-5) on the shipped checkerboard points of sunday_amelia -
    the runs whose RMS the reference's notebook prints (calib_with_gui.ipynb:665,673: 0.32182 / 0.36876 px)."""
    import cv2

    base = os.path.join(REF, "data", "sunday_amelia", "extrinsic_calib")
    out = {}
    for tag, scene, cams in [("rot12", "4_cam_scene_rotating.json", (1, 2)), ("sta34", "4_cam_scene_static.json", (3, 4))]:
        K, D, R, t, res = utils.load_scene(os.path.join(base, scene))
        pts, fns = [], []
        for c in cams:
            d = json.load(open(os.path.join(base, "points", f"points_cam{c}.json")))
            fns.append(list(d["points"].keys()))
            pts.append(np.array(list(d["points"].values()), dtype=np.float32))
            board_shape, bel = tuple(d["board_shape"]), d.get("board_edge_len", d.get("board_square_len"))
        common = [f for f in fns[0] if f in fns[1]]
        p1 = np.array([pts[0][fns[0].index(f)] for f in common], dtype=np.float32)
        p2 = np.array([pts[1][fns[1].index(f)] for f in common], dtype=np.float32)
        obj = utils.create_board_object_pts(board_shape, bel)
        n = len(common)
        ia, ib = (cams[0] - 1, cams[1] - 1) if len(K) == 4 else (0, 1)
        objp = np.repeat(obj[np.newaxis], n, axis=0).reshape(n, 1, -1, 3)
        res_cv = cv2.fisheye.stereoCalibrate(objp, p1.reshape(n, 1, -1, 2), p2.reshape(n, 1, -1, 2), K[ia].copy(), D[ia].copy(),
                                             K[ib].copy(), D[ib].copy(), res, flags=cv2.fisheye.CALIB_FIX_INTRINSIC,
                                             criteria=(cv2.TERM_CRITERIA_MAX_ITER + cv2.TERM_CRITERIA_EPS, 100, 1e-5))
        out.update({f"{tag}_obj": obj, f"{tag}_img1": p1, f"{tag}_img2": p2, f"{tag}_K1": K[ia], f"{tag}_D1": D[ia],
                    f"{tag}_K2": K[ib], f"{tag}_D2": D[ib], f"{tag}_rms": res_cv[0], f"{tag}_R": res_cv[5], f"{tag}_T": res_cv[6],
                    f"{tag}_scene_R": R[[ia, ib]], f"{tag}_scene_t": t[[ia, ib]], f"{tag}_res": np.array(res)})
        print("stereo", tag, "views", n, "rms", res_cv[0])
    # standard camera model: the reference's calibrate_pair_extrinsics (calib.py:41-49), run unmodified, on synthetic boards
    rng = np.random.default_rng(17)
    PIN_NOISE = 0.0      # cv2.stereoCalibrate stops early (30 iterations, eps 1e-5): with noise it ends far from the minimum
    K1 = np.array([[1210.0, 0, 960.0], [0, 1198.0, 540.0], [0, 0, 1]])
    K2 = np.array([[1180.0, 0, 950.0], [0, 1185.0, 545.0], [0, 0, 1]])
    d1 = np.array([0.11, -0.05, 0.001, -0.002, 0.01, 0.02, -0.01, 0.003])
    d2 = np.array([-0.08, 0.03, -0.0007, 0.0012, 0.004, 0.015, -0.006, 0.001])
    obj = utils.create_board_object_pts((9, 6), 0.05).astype(np.float32)
    R_true = cv2.Rodrigues(np.array([0.04, -0.35, 0.02]))[0]
    T_true = np.array([-0.55, 0.02, 0.08])
    V = 14
    i1, i2 = np.empty((V, 54, 2), np.float32), np.empty((V, 54, 2), np.float32)
    for v in range(V):
        Rv = cv2.Rodrigues(rng.normal(0, 0.25, 3))[0]
        tv = np.array([rng.uniform(-0.25, 0.25), rng.uniform(-0.15, 0.15), rng.uniform(0.9, 1.5)])
        u1 = calib.project_points(obj.astype(np.float64), K1, d1, Rv, tv)
        u2 = calib.project_points(obj.astype(np.float64), K2, d2, R_true @ Rv, R_true @ tv + T_true)
        i1[v] = u1 + rng.normal(0, PIN_NOISE, u1.shape)
        i2[v] = u2 + rng.normal(0, PIN_NOISE, u2.shape)
    rms_p, r_p, t_p = calib.calibrate_pair_extrinsics(obj, i1.reshape(V, 9, 6, 2), i2.reshape(V, 9, 6, 2), K1, d1, K2, d2, (1920, 1080))
    out.update(pin_obj=obj, pin_img1=i1, pin_img2=i2, pin_K1=K1, pin_D1=d1, pin_K2=K2, pin_D2=d2, pin_rms=rms_p, pin_R=r_p,
               pin_T=t_p, pin_R_true=R_true, pin_T_true=T_true)
    print("stereo pinhole: rms", rms_p, "dR", np.abs(r_p - R_true).max(), "dT", np.abs(t_p.ravel() - T_true).max())
    out["notebook_rms"] = np.array([0.32182, 0.36876])
    np.savez_compressed(os.path.join(HERE, "stereo.npz"), **out)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-sympy-jac", action="store_true")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    calib, utils = ref_shim.load_reference_calib()
    todo = args.only.split(",") if args.only else ["fk", "fisheye", "loss", "tri", "generic", "sba", "pinhole", "stereo"]
    if "fisheye" in todo:
        gen_fisheye(calib, utils)
    if "loss" in todo:
        gen_loss()
    if "tri" in todo:
        gen_triangulate(calib, utils)
    if "generic" in todo:
        gen_generic_fk()
    if "sba" in todo:
        gen_sba(calib, utils)
    if "pinhole" in todo:
        gen_pinhole(calib, utils)
    if "stereo" in todo:
        gen_stereo(calib, utils)
    if "fk" in todo:
        gen_cheetah_fk(args.skip_sympy_jac)


if __name__ == "__main__":
    main()
