"""Injected-draw streams shared by the golden generator and the tests (test infrastructure)."""
import numpy as np

NORMAL_SLOTS = (2, 7, 10, 11)
N_SLOTS, MAX_ITERS = 16, 360


def adversarial_stream(n_games: int, seed: int = 2025) -> np.ndarray:
    """[n, 360, 16] float64 draw records that force the rare branches of the engine: game g uses pattern
    g % 6 -- (0) every uniform 0, normals -4; (1) uniforms just below 1, normals +4; (2) u = .5, z = 0;
    (3) ordinary draws with 35 % of the slots pushed to an extreme; (4) u = 1e-12, z = 3; (5) u = .999, z = -3."""
    rng = np.random.default_rng(seed)
    s = rng.random((n_games, MAX_ITERS, N_SLOTS))
    s[:, :, NORMAL_SLOTS] = rng.standard_normal((n_games, MAX_ITERS, len(NORMAL_SLOTS)))
    uni = [k for k in range(N_SLOTS) if k not in NORMAL_SLOTS]
    const = {0: (0.0, -4.0), 1: (1.0 - 2.0 ** -40, 4.0), 2: (0.5, 0.0), 4: (1e-12, 3.0), 5: (0.999, -3.0)}
    for g in range(n_games):
        k = g % 6
        if k in const:
            s[g][:, uni] = const[k][0]
            s[g][:, NORMAL_SLOTS] = const[k][1]
        else:
            mask = rng.random(s[g].shape) < 0.35
            ext = np.where(rng.random(s[g].shape) < 0.5, 0.0, 1.0 - 2.0 ** -33)
            zext = np.where(rng.random(s[g].shape) < 0.5, -5.0, 5.0)
            sg = s[g]
            sg[:, uni] = np.where(mask[:, uni], ext[:, uni], sg[:, uni])
            sg[:, NORMAL_SLOTS] = np.where(mask[:, NORMAL_SLOTS], zext[:, NORMAL_SLOTS], sg[:, NORMAL_SLOTS])
    return s
