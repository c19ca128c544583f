import sys, numpy as np
sys.path.insert(0, '.')
import synth
from oracle import fisheye, skeleton, fte, loss
import acinoset_b200 as ab
cams = synth.load_dummy_scene()
K, D, R, t, _ = cams
h = ab.Handle(0); h.set_cameras(K, D, R, t)
p = synth.make_fte_problem(1000, skeleton.cheetah_fk_active, fisheye.project, seed=5, cams=cams)
x32 = p["x0"].astype(np.float32); m32 = p["meas"].astype(np.float32); w32 = p["w"].astype(np.float32)
cost, g, Hu = h.fte_eval(x32, m32, w32)
c_ref, g_ref, H_ref = fte.fte_eval(x32.astype(np.float64), m32.astype(np.float64), w32.astype(np.float64), K, D, R, t)
np.set_printoptions(linewidth=200, precision=4, suppress=False)
H = fte.unpack_upper(Hu.astype(np.float64))
herr = np.linalg.norm(H - H_ref, axis=(1, 2)) / np.linalg.norm(H_ref, axis=(1, 2))
print("herr sorted top", np.sort(herr)[-5:], "median", np.median(herr))
f = int(np.argmax(herr))
print("frame", f, "|H|", np.linalg.norm(H_ref[f]), "nvalid", (w32[f]>0).sum())
r, J = fte.residuals_and_jac(x32[f:f+1].astype(np.float64), m32[f:f+1].astype(np.float64), K, D, R, t)
for c in range(6):
    for l in range(20):
        if w32[f,c,l] == 0: continue
        w1 = np.zeros_like(w32[f:f+1]); w1[0,c,l] = w32[f,c,l]
        _, g1, H1 = h.fte_eval(x32[f:f+1], m32[f:f+1], w1)
        _, g1r, H1r = fte.fte_eval(x32[f:f+1].astype(np.float64), m32[f:f+1].astype(np.float64), w1.astype(np.float64), K, D, R, t)
        dH = np.linalg.norm(fte.unpack_upper(H1.astype(np.float64))-H1r)
        if dH > 1e-5*np.linalg.norm(H_ref[f]):
            e = 0.2*r[0,c,l]
            print("cam", c, "marker", l, "abs dH", dH, "|H1|", np.linalg.norm(H1r), "e", e, "psi", loss.redescending_irls_weight(e), "|J|", np.linalg.norm(J[0,c,l],axis=1))
