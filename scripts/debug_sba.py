import sys, numpy as np
sys.path.insert(0, '.')
from acinoset_b200 import calib, sba
g = np.load('tests/golden/sba.npz')
tag = sys.argv[1] if len(sys.argv) > 1 else 'static'
K, D, R, t = g[f"{tag}_K"], g[f"{tag}_D"], g[f"{tag}_R"], g[f"{tag}_t"]
obj, r_new, t_new, res, info = sba.bundle_adjust_points_and_extrinsics(
    g[f"{tag}_points_2d"], g[f"{tag}_points_3d"], g[f"{tag}_pidx"], g[f"{tag}_cidx"], K, D, R, t, None, verbose=2, return_info=True)
print({k: v for k, v in info.items() if k not in ('params', 'pts', 'fun', 'f0')})
