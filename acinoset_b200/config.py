"""The constants of the two FTE formulations in one place ("config" row): the reference scatters
them through its scripts (all_optimizations.py:243-252,403-483,497,511; build.py:131-135,299,312); the modules here
read them from these frozen dataclasses so that a caller can see - and, by passing a replaced copy where an entry
point takes one, change - every number the solve depends on."""
from dataclasses import dataclass, field

import numpy as np

_P6, _P15, _P2, _PI = np.pi / 6, np.pi / 1.5, np.pi / 2, np.pi


def _frozen(a):
    a = np.asarray(a, dtype=np.float64)
    a.setflags(write=False)
    return a


def _cheetah_bounds():
    """(lo, hi)[25] of all_optimizations.py:403-483 in the active ordering; +-inf where free."""
    lo = np.full(25, -np.inf)
    hi = np.full(25, np.inf)
    for i in (3, 4, 5, 6, 7, 8, 9, 21, 22):          # phi0 phi1 phi3 theta0..3 psi1 psi3
        lo[i], hi[i] = -_P6, _P6
    for i in (10, 11, 23, 24):                        # theta4 theta5 psi4 psi5
        lo[i], hi[i] = -_P15, _P15
    for i in (12, 14, 16, 18):                        # shoulders / hips
        lo[i], hi[i] = -_P2, _P2
    for i in (13, 15):                                # front knees
        lo[i], hi[i] = -_PI, 0.0
    for i in (17, 19):                                # back knees
        lo[i], hi[i] = 0.0, _PI
    return _frozen(lo), _frozen(hi)


@dataclass(frozen=True)
class CheetahFTEConfig:
    """all_optimizations.py (the cheetah pipeline script)."""
    meas_sigma_px: float = 5.0                        # R, :243 - measurement weight 1/R where likelihood > dlc_thresh
    dlc_thresh: float = 0.5                           # default of the script's --dlc_thresh argument
    redescending: tuple = (3.0, 10.0, 20.0)           # break points a, b, c of redescending_loss, :497
    # model (acceleration-slack) std-devs Q of the 25 active slots, :245-252; weight 1/Q^2
    q_sigma: np.ndarray = field(default_factory=lambda: _frozen(
        [4, 7, 5, 13, 32, 10, 9, 18, 43, 53, 90, 118, 247, 186, 194, 164, 295, 243, 334, 149, 26, 12, 34, 43, 51]))
    bounds: tuple = field(default_factory=_cheetah_bounds)     # (lo, hi) of the 21 bounded slots, :403-483
    fps_default: float = 90.0                         # Ts = 1 / fps when the video meta data gives nothing else
    # projected Levenberg-Marquardt loop that replaces IPOPT (tol 1e-1, max_iter 10 000 at :511-514)
    lm_lam0: float = 1e-3
    lm_tol_step: float = 1e-6
    lm_tol_rel: float = 1e-7
    lm_tol_noise: float = 5e-8
    lm_max_attempts: int = 12
    lm_max_iter: int = 60


@dataclass(frozen=True)
class SkeletonFTEConfig:
    """build.py (the generic-skeleton formulation)."""
    model_weight: float = 0.002                       # build.py:312: 0.002 * slack_model^2 (instead of 1/Q^2)
    meas_sigma_px: float = 3.0                        # R, build.py:131
    lik_thresh: float = 0.4                           # build.py:189-198
    redescending: tuple = (3.0, 10.0, 20.0)
    abs_delta: float = 0.05                           # curvature floor of the |w r| objective (build.py:299 has no curvature)
    angle_bound: float = float(np.pi / 2)             # build.py:238-247
    n_frames: int = 100                               # build.py:131-135 (frames 60..159)


CHEETAH = CheetahFTEConfig()
SKELETON = SkeletonFTEConfig()
